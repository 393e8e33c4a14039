"""`FeedingEnv.step` semantics (reference envs/feeding.py:12-112 + env.py:174-274): the repo's numpy restatement
(`tests/parity_cases.feeding_semantics_reference` + `take_step_targets`, the functions the fused CUDA kernels are checked against
in tests/test_gpu_parity.py / test_kernel_logic_cpu.py) replays the rollout of tests/golden/feeding_semantics.npz, which was
produced by the reference's OWN step code running on the CPU oracle through a pybullet facade
(tests/golden/make_golden_feeding_semantics.py).  Same physics under both (the oracle), so observation, reward, done and the food
bookkeeping must agree to rounding: a spilled particle (-5), an eaten one (+20, its speed penalised) and 38 ordinary steps."""
import os

import numpy as np

from assistive_gym_b200 import capi
from assistive_gym_b200.feeding_batch import FeedingBatch
from oracle.oracle_py import OracleSim
from tests.parity_cases import feeding_semantics_reference, take_step_targets

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'feeding_semantics.npz'))


def test_restated_feeding_step_reproduces_the_reference_s_rollout():
    fb = FeedingBatch()
    sim = OracleSim(fb.scene, capi.default_config(), 1)
    smp = {k[len('sample_'):]: G[k] for k in G.files if k.startswith('sample_')}
    fb.reset(sim, np.random.default_rng(int(G['seed'])), settle_steps=25, impairment='none')       # the generator's call: same draws, same IK restarts
    assert all(np.array_equal(np.asarray(fb.last_sample[k]), smp[k]) for k in smp if k in fb.last_sample)
    sim.state_set(G['start_state']); sim.forward_kinematics()               # exactly the generator's start state (the reset's IK goes through BLAS)
    state = dict(male=smp['male'], foods=np.ones((1, 8), dtype=bool), active=np.ones((1, 8), dtype=bool), iteration=np.zeros(1, dtype=int), task_success=np.zeros(1, dtype=int))
    rng_far = np.random.RandomState(0)
    events = []
    for t, a in enumerate(G['actions']):
        if t == int(G['eat_step']):
            # the forced event of the golden rollout: a particle tossed up from the mouth target
            ls = sim.get_link_states([fb.gl(fb.humans['male' if smp['male'][0] else 'female'], 23)])
            target = ls['pos'][0, 0] + _qrot(ls['quat'][0, 0], fb.mouth['male' if smp['male'][0] else 'female'])
            f = fb.foods[int(G['eat_food'])]
            sim.set_base_pose(f, target[None], np.array([[0, 0, 0, 1.0]]))
            sim.set_base_velocity(f, np.array([[0, 0, float(G['eat_v0'])]]), np.zeros((1, 3)))
        q = sim.get_joint_states(fb.arm_links)[0]
        sim.set_motor_targets(fb.arm_links, take_step_targets(q, a[None], fb.arm_lower, fb.arm_upper))
        sim.step(5)
        obs, rew, done, total = feeding_semantics_reference(fb, sim, a[None], state)
        assert np.allclose(obs[0], G['obs'][t], rtol=0, atol=1e-9), (t, np.abs(obs[0] - G['obs'][t]).max())
        assert abs(rew[0] - G['reward'][t]) < 1e-9, (t, rew[0], G['reward'][t])
        assert bool(done[0]) == bool(G['done'][t]) and abs(total[0] - G['total_force'][t]) < 1e-9
        assert int(state['foods'].sum()) == int(G['n_foods'][t]) and int(state['active'].sum()) == int(G['n_foods_active'][t])
        assert int(state['task_success'][0]) == int(G['task_success'][t])
        for i in np.where(state['eaten_now'][0])[0]:                    # feeding.py:69: an eaten particle is moved far away
            sim.set_base_pose(fb.foods[i], rng_far.uniform(1000, 2000, size=(1, 3)), np.array([[0, 0, 0, 1.0]]))
            events.append(('eaten', t))
        if G['reward'][t] < -4:
            events.append(('spilled', t))
    assert ('eaten', int(G['eat_step'])) in events and any(e[0] == 'spilled' for e in events)


def _qrot(q, v):
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return R @ np.asarray(v, dtype=np.float64)


def test_fused_kernel_bodies_reproduce_the_reference_s_rollout(emu_lib):
    """The product's fused Feeding step (`ag_feeding_step_host`: the CUDA kernel bodies, here compiled for the host, fp32) driven with the
    golden rollout's start state and actions, against what the reference's own `FeedingEnv.step` returned on the fp64 oracle: no
    restatement in between.  North-star tolerances (1e-3 m, 1e-4 rad) over the first env steps; after the spilled particle (step 4)
    the two physics paths drift apart at the 1e-4 level, which the bounds below allow for.  Rewards agree incl. the -5 of the spill
    and the +20 of the particle tossed into the mouth."""
    from assistive_gym_b200.sim import BatchSim
    fb = FeedingBatch()
    prod = BatchSim(fb.scene, capi.default_config(), 1, _lib=emu_lib)
    smp = {k[len('sample_'):]: G[k] for k in G.files if k.startswith('sample_')}
    fb.reset(prod, np.random.default_rng(int(G['seed'])), settle_steps=0, sample=smp)
    prod.state_set(G['start_state'].astype(np.float32)); prod.forward_kinematics()
    fb.start_fused(prod, smp, seed=1)
    male = bool(smp['male'][0])
    for t, a in enumerate(G['actions'][:20]):
        if t == int(G['eat_step']):
            ls = prod.get_link_states([fb.gl(fb.humans['male' if male else 'female'], 23)])
            target = ls['pos'][0, 0].astype(np.float64) + _qrot(ls['quat'][0, 0].astype(np.float64), fb.mouth['male' if male else 'female'])
            f = fb.foods[int(G['eat_food'])]
            prod.set_base_pose(f, target[None], np.array([[0, 0, 0, 1.0]]))
            prod.set_base_velocity(f, np.array([[0, 0, float(G['eat_v0'])]]), np.zeros((1, 3)))
        obs, rew, done, info = prod.feeding_step_host(a[None].astype(np.float32))
        e = np.abs(obs[0] - G['obs'][t])
        tight = t <= 4
        assert e[[0, 1, 2, 7, 8, 9, 17, 18, 19]].max() < (5e-6 if tight else 1e-3), (t, e)          # positions (spoon, spoon - target, head), metres
        assert e[10:17].max() < (5e-6 if tight else 1e-3) and e[3:7].max() < (5e-6 if tight else 1e-3)      # joint angles (rad), spoon orientation
        assert abs(rew[0] - G['reward'][t]) < (1e-5 if tight else 2e-2), (t, rew[0], G['reward'][t])
    assert G['reward'][4] < -5 and G['reward'][int(G['eat_step'])] > 18
