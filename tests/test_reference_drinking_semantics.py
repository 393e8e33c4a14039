"""`DrinkingEnv.step` semantics (reference envs/drinking.py:10-117 + env.py:174-274): `DrinkingJacoEnv` of this repo (per-call API path,
vectorised water bookkeeping), run on the CPU oracle, replays the rollout of tests/golden/drinking_semantics.npz, produced by the
reference's OWN step code on the same oracle through a pybullet facade (tests/golden/make_golden_drinking_semantics.py): the cup with
64 water particles at 4 substeps / 10 solver iterations, one particle swallowed (+10, its speed penalised), one spilled (-1)."""
import os

import numpy as np

from assistive_gym_b200 import envs
from assistive_gym_b200.drinking_batch import DrinkingBatch
from oracle.oracle_py import OracleSim

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'drinking_semantics.npz'))


def test_drinking_step_reproduces_the_reference_s_rollout():
    db = DrinkingBatch()
    sim = OracleSim(db.scene, DrinkingBatch.config(), 1)
    smp = db.reset(sim, np.random.default_rng(int(G['seed'])), settle_steps=50, impairment='none')                 # the generator's call
    sim.state_set(G['start_state']); sim.forward_kinematics()
    env = envs.make('DrinkingJaco-v1', n_envs=1)
    env._db = db
    env.attach(sim)
    env.start_episode(smp)
    for t, a in enumerate(G['actions']):
        if t == int(G['swallow_step']):
            w = db.waters[int(G['swallow_water'])]
            sim.set_base_pose(w, G['swallow_pos'][None], np.array([[0, 0, 0, 1.0]]))
            sim.set_base_velocity(w, np.array([[0, 0, float(G['swallow_v0'])]]), np.zeros((1, 3)))
        if t == int(G['spill_step']):
            w = db.waters[int(G['spill_water'])]
            sim.set_base_pose(w, G['spill_pos'][None], np.array([[0, 0, 0, 1.0]]))
            sim.set_base_velocity(w, np.zeros((1, 3)), np.zeros((1, 3)))
        obs, rew, done, info = env.step(a)
        assert np.allclose(obs[:24], G['obs'][t][:24], rtol=0, atol=1e-6), (t, np.abs(obs - G['obs'][t]).max())
        assert abs(obs[24] - G['obs'][t][24]) < 1e-4 * (1 + abs(G['obs'][t][24])) and abs(rew - G['reward'][t]) < 1e-5, (t, rew, G['reward'][t])
        assert bool(done) == bool(G['done'][t]) and abs(info['total_force_on_human'] - G['total_force'][t]) < 1e-4 * (1 + G['total_force'][t])
        assert int(env.waters.sum()) == int(G['n_waters'][t]) and int(env.waters_active.sum()) == int(G['n_waters_active'][t])
        assert int(env.task_success[0]) == int(G['task_success'][t])
    assert G['task_success'][-1] == 1 and G['reward'][int(G['swallow_step'])] > 8 and G['reward'][int(G['spill_step'])] < -1.5


def test_drinking_env_on_the_host_compiled_kernel_bodies(emu_lib):
    """the product path (kernel bodies compiled for the host): reset, the water settles in the cup, a few steps run and stay finite"""
    env = envs.make('DrinkingJaco-v1', n_envs=2, seed=3)
    env._sim_lib = emu_lib
    obs = env.reset()
    assert obs.shape == (2, 25) and np.all(np.isfinite(obs)) and env.action_space.shape == (7,)
    assert env.id.overflow_count() == 0
    top, bottom, _ = env._cup_centres()
    from assistive_gym_b200.envs.drinking import points_in_cylinder
    sc = env.id.scene
    wp = env.id.get_link_states([int(sc['body_link0'][w.body]) for w in env.water_agents])['pos'].astype(np.float64)
    assert points_in_cylinder(top, bottom, 0.05, wp).sum(axis=1).min() >= 56          # the water is in the cup after the 50 settle steps
    assert env._db.ik_colliding == 0                                              # start poses that touch the person / wheelchair are resampled (env.py:300-309)
    for _ in range(2):
        o, r, d, info = env.step(np.zeros((2, 7)))
        assert np.all(np.isfinite(o)) and np.all(np.isfinite(r))
    assert env.waters.sum(axis=1).min() >= 56
    env.close()


def test_drinking_env_on_the_kernel_bodies_reproduces_the_reference_s_rollout(emu_lib):
    """`DrinkingJacoEnv` on the PRODUCT's physics (kernel bodies compiled for the host, fp32; 65 free bodies per env, 4 substeps, 10 solver
    iterations) from the golden start, against what the reference's own `DrinkingEnv.step` returned on the fp64 oracle: observation 1e-4,
    reward 1e-2 incl. the swallowed (+10) and the spilled (-1) particle, the same particles removed."""
    from assistive_gym_b200.sim import BatchSim
    db = DrinkingBatch()
    prod = BatchSim(db.scene, DrinkingBatch.config(), 1, _lib=emu_lib)
    smp = {k[len('sample_'):]: G[k] for k in G.files if k.startswith('sample_')}
    db.reset(prod, np.random.default_rng(0), settle_steps=0, sample=smp)
    prod.state_set(G['start_state'].astype(np.float32)); prod.forward_kinematics()
    env = envs.make('DrinkingJaco-v1', n_envs=1)
    env._db = db
    env.attach(prod)
    env.start_episode(smp)
    for t, a in enumerate(G['actions'][:16]):
        for key in ('swallow', 'spill'):
            if t == int(G[key + '_step']):
                w = db.waters[int(G[key + '_water'])]
                prod.set_base_pose(w, G[key + '_pos'][None], np.array([[0, 0, 0, 1.0]]))
                prod.set_base_velocity(w, np.array([[0, 0, float(G['swallow_v0']) if key == 'swallow' else 0.0]]), np.zeros((1, 3)))
        obs, rew, done, info = env.step(a)
        assert np.abs(obs[:24] - G['obs'][t][:24]).max() < 1e-4 and abs(rew - G['reward'][t]) < 1e-2, (t, rew, G['reward'][t])
        assert int(env.waters.sum()) == int(G['n_waters'][t]) and int(env.task_success[0]) == int(G['task_success'][t])
    assert prod.overflow_count() == 0
