"""Co-optimisation step of ScratchItch (reference envs/scratch_itch.py:10-91 with dict actions, env.py:174-235 incl.
`Human.enforce_realistic_joint_limits`, agents/human.py:134-152): `ScratchItchJacoHumanEnv` of this repo, run on the CPU oracle,
replays the rollout of tests/golden/scratch_coop_semantics.npz, produced by the reference's OWN step code on the same oracle
through a pybullet facade (tests/golden/make_golden_scratch_coop_semantics.py).  The person raises the upper arm until the
joint-limit classifier stops it at 116.5 degrees; both dict observations, the reward and the arm's joint angles must agree."""
import os

import numpy as np

from assistive_gym_b200 import capi, envs
from assistive_gym_b200.scratch_itch_batch import RIGHT_ARM_JOINTS, ScratchItchBatch
from oracle.oracle_py import OracleSim

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'scratch_coop_semantics.npz'))


def test_cooptimisation_step_reproduces_the_reference_s_rollout():
    sb = ScratchItchBatch()
    sim = OracleSim(sb.scene, capi.default_config(), 1)
    smp = {k[len('sample_'):]: G[k] for k in G.files if k.startswith('sample_')}
    sb.reset(sim, np.random.default_rng(0), sample=smp)
    env = envs.make('ScratchItchJacoHuman-v1', n_envs=1)
    env._sb = sb
    env.id = sim                                                               # the env's per-call path on the oracle instead of the CUDA library
    env.plane.init(sb.plane, sim, env.np_random, indices=-1)
    env.robot.init(sb.robot, sim, env.np_random)
    env.tool.init(sb.tool, sim, env.np_random, indices=-1)
    env.furniture.init(sb.wheelchair, sim, env.np_random, indices=-1)
    env.male = smp['male'].astype(bool)
    env.humans = {}
    env.agents = [env.robot]
    for g, hb in sb.humans.items():
        h = type(env.human)(env.human.controllable_joint_indices, controllable=True)
        h.init(hb, sim, env.np_random, env.human.controllable_joint_indices)
        h.env_mask = env.male if g == 'male' else ~env.male
        h.set_limit_scale(np.ones(1))
        env.humans[g] = h
        env.agents.append(h)
    env._limb_links, env._target_local = sb.limb_links(smp), smp['target_local']
    env.prev_target_contact_pos = np.zeros((1, 3))
    env.task_success = np.zeros(1, dtype=int)
    env.iteration = 0
    hb = sb.humans['male' if env.male[0] else 'female']
    links = [sb.gl(hb, j) for j in RIGHT_ARM_JOINTS]
    for t in range(len(G['reward'])):
        o, r, d, info = env.step({'robot': np.zeros(7), 'human': G['human_action']})
        arm = sim.get_joint_states(links)[0][0]
        assert np.allclose(arm, G['arm_q'][t], rtol=0, atol=1e-7), (t, np.abs(arm - G['arm_q'][t]).max())
        assert np.allclose(o['robot'][:29], G['obs_robot'][t][:29], rtol=0, atol=1e-6) and abs(o['robot'][29] - G['obs_robot'][t][29]) < 1e-4 * (1 + abs(G['obs_robot'][t][29]))
        assert np.allclose(o['human'][:32], G['obs_human'][t][:32], rtol=0, atol=1e-6) and np.allclose(o['human'][32:], G['obs_human'][t][32:], rtol=1e-4, atol=1e-4)
        assert abs(r['robot'] - G['reward'][t]) < 1e-5 and r['robot'] == r['human']
    assert abs(np.rad2deg(G['arm_q'][-1, 3]) - 116.5) < 0.5 and np.rad2deg(G['arm_q'][20, 3]) < 100          # the classifier, not the joint range (198 degrees), ends the motion
