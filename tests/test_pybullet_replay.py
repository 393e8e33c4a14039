"""Pin to the real engine (VERDICT r1 missing #4): replay of the reference's own FeedingJaco-v1 in
PyBullet against this backend -- same URDFs, same actions, the three north-star numbers (1e-4 rad, 1e-3 m over 200
substeps, 5 % tool-on-body force).

SKIPPED wherever PyBullet, gym or the reference package are not importable -- which includes the build container and
the GPU box of this project (no network, `pybullet` is not in the wheelhouse): this file has never been executed and
the oracle therefore stays "parity unpinned" (DESIGN.md section 5).  It documents the protocol a maintainer with
PyBullet runs: `AG_REFERENCE_PATH=/path/to/assistive-gym pytest tests/test_pybullet_replay.py`."""
import os
import sys

import numpy as np
import pytest

pybullet = pytest.importorskip('pybullet')
gym = pytest.importorskip('gym')


def _reference_env():
    ref = os.environ.get('AG_REFERENCE_PATH', '/root/reference')
    if not os.path.isdir(os.path.join(ref, 'assistive_gym')):
        pytest.skip('reference package not found (set AG_REFERENCE_PATH)')
    sys.path.insert(0, ref)
    for m in [m for m in sys.modules if m == 'assistive_gym' or m.startswith('assistive_gym.')]:
        del sys.modules[m]                      # this repo ships a drop-in shim of the same name
    import assistive_gym  # noqa: F401
    if getattr(assistive_gym, '__agphys_shim__', False):
        pytest.skip('the shim shadows the reference package')
    env = gym.make('assistive_gym:FeedingJaco-v1')
    env.seed(1001)
    return env


@pytest.mark.gpu
def test_feeding_jaco_replay_against_pybullet():
    import pybullet as p
    from assistive_gym_b200 import capi
    from assistive_gym_b200.feeding_batch import FeedingBatch
    from assistive_gym_b200.sim import BatchSim
    env = _reference_env()
    env.reset()
    ref = env.unwrapped
    fb = FeedingBatch()
    sim = BatchSim(fb.scene, capi.default_config(), 1, device=0)
    s = fb.sample(1, np.random.default_rng(0))
    s['male'][:] = 1 if ref.human.gender == 'male' else 0
    s['impairment'][:] = 0
    fb.reset(sim, np.random.default_rng(0), settle_steps=0, sample=s)
    # copy the reference's start state: arm + gripper angles, robot base, spoon, bowl, food, head angles
    arm = ref.robot.controllable_joint_indices
    q_arm = np.array(ref.robot.get_joint_angles(arm))[None]
    sim.set_joint_state(fb.arm_links, q_arm, np.zeros_like(q_arm))
    sim.set_motor_targets(fb.arm_links, q_arm)
    for body, agent in ((fb.tool, ref.tool), (fb.bowl, ref.furniture if hasattr(ref, 'bowl') else ref.bowl)):
        pos, orn = agent.get_base_pos_orient()
        sim.set_base_pose(body, np.array(pos)[None], np.array(orn)[None])
    for f, food in zip(fb.foods, ref.foods):
        pos, orn = food.get_base_pos_orient()
        sim.set_base_pose(f, np.array(pos)[None], np.array(orn)[None])
    sim.forward_kinematics()
    rng = np.random.default_rng(0)
    err_q = err_tool = 0.0
    forces = []
    from tests.parity_cases import take_step_targets
    for _ in range(40):                          # 40 env steps x 5 = 200 substeps
        a = rng.uniform(-1, 1, size=7)
        env.step(a)
        tgt = take_step_targets(sim.get_joint_states(fb.arm_links)[0], a[None], fb.arm_lower, fb.arm_upper)
        sim.set_motor_targets(fb.arm_links, tgt)
        sim.step(5)
        err_q = max(err_q, np.abs(np.array(ref.robot.get_joint_angles(arm)) - sim.get_joint_states(fb.arm_links)[0][0]).max())
        tl = int(fb.scene['body_link0'][fb.tool])
        err_tool = max(err_tool, np.abs(np.array(ref.tool.get_base_pos_orient()[0]) - sim.get_link_states([tl])['pos'][0, 0]).max())
        f_ref = sum(c[9] for c in p.getContactPoints(bodyA=ref.tool.body, bodyB=ref.human.body, physicsClientId=ref.id))
        f_dev = float(sim.contact_force_sum(fb.tool, fb.humans['male' if s['male'][0] else 'female'])[0])
        if f_ref > 0.5:
            forces.append(abs(f_ref - f_dev) / f_ref)
    print('pybullet replay: |dq| %.3g rad, |dtool| %.3g m, tool-on-body force rel %s' % (err_q, err_tool, max(forces) if forces else None))
    assert err_q < 1e-4 and err_tool < 1e-3
    assert not forces or max(forces) < 0.05
