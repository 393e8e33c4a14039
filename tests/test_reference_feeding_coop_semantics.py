"""Co-optimisation step of Feeding (reference envs/feeding.py:12-112 with dict actions and the person's observation :101-111,
env.py:174-235 with the person's head joints as a second agent): `FeedingJacoHumanEnv` of this repo, run on the CPU oracle,
replays the rollout of tests/golden/feeding_coop_semantics.npz, produced by the reference's OWN step code on the same oracle
through a pybullet facade (tests/golden/make_golden_feeding_coop_semantics.py)."""
import os

import numpy as np

from assistive_gym_b200 import capi, envs
from assistive_gym_b200.envs.agents.agent import Agent
from assistive_gym_b200.envs.agents.furniture import Furniture
from assistive_gym_b200.feeding_batch import FeedingBatch
from oracle.oracle_py import OracleSim

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'feeding_coop_semantics.npz'))
HEAD = [20, 21, 22, 23]


def test_cooptimisation_feeding_step_reproduces_the_reference_s_rollout():
    fb = FeedingBatch()
    sim = OracleSim(fb.scene, capi.default_config(), 1)
    smp = fb.reset(sim, np.random.default_rng(int(G['seed'])), settle_steps=25, impairment='none', simulate_head=True)       # the generator's call
    assert all(np.array_equal(np.asarray(smp[k]), G['sample_' + k]) for k in smp if 'sample_' + k in G.files)
    sim.state_set(G['start_state']); sim.forward_kinematics()               # exactly the generator's start state
    env = envs.make('FeedingJacoHuman-v1', n_envs=1)
    env._fb = fb
    env.id = sim                                                               # the env's per-call path on the oracle instead of the CUDA library
    env.plane.init(fb.plane, sim, env.np_random, indices=-1)
    env.robot.init(fb.robot, sim, env.np_random)
    env.tool.init(fb.tool, sim, env.np_random, indices=-1)
    env.furniture.init(fb.wheelchair, sim, env.np_random, indices=-1)
    env.table, env.bowl = Furniture(), Furniture()
    env.table.init(fb.table, sim, env.np_random, indices=-1)
    env.bowl.init(fb.bowl, sim, env.np_random, indices=-1)
    env.male = smp['male'].astype(bool)
    env.robot.motor_gains = env.human.motor_gains = 0.025
    env.humans, env.agents = {}, [env.robot]
    for g, hb in fb.humans.items():
        h = type(env.human)(env.human.controllable_joint_indices, controllable=True)
        h.init(hb, sim, env.np_random, env.human.controllable_joint_indices)
        h.motor_gains, h.motor_forces = 0.025, 1.0
        h.tremor_mask = np.zeros(1, dtype=bool)
        env.humans[g] = h
        env.agents.append(h)
    env.foods_agents = []
    for f in fb.foods:
        a = Agent()
        a.init(f, sim, env.np_random, indices=-1)
        env.foods_agents.append(a)
    env.mouth_pos = np.where(env.male[:, None], fb.mouth['male'], fb.mouth['female'])
    env.foods = np.ones((1, 8), dtype=bool); env.foods_active = np.ones((1, 8), dtype=bool)
    env.task_success = np.zeros(1, dtype=int)
    env.iteration = 0
    env.update_targets()
    hb = fb.humans['male' if env.male[0] else 'female']
    links = [fb.gl(hb, j) for j in HEAD]
    for t in range(len(G['reward'])):
        o, r, d, info = env.step({'robot': G['robot_actions'][t], 'human': G['human_actions'][t]})
        head = sim.get_joint_states(links)[0][0]
        assert np.allclose(head, G['head_q'][t], rtol=0, atol=1e-7), (t, np.abs(head - G['head_q'][t]).max())
        assert np.allclose(o['robot'], G['obs_robot'][t], rtol=0, atol=1e-6), (t, np.abs(o['robot'] - G['obs_robot'][t]).max())
        assert np.allclose(o['human'], G['obs_human'][t], rtol=0, atol=1e-6), (t, np.abs(o['human'] - G['obs_human'][t]).max())
        assert abs(r['robot'] - G['reward'][t]) < 1e-5 and r['robot'] == r['human']
    assert np.all(G['head_q'] >= G['head_lower'] - 1e-9) and np.all(G['head_q'] <= G['head_upper'] + 1e-9)
    assert np.abs(G['head_q'][11] - G['head_q'][0]).max() > 0.2               # the head followed the person's action
