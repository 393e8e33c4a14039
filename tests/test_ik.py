"""Batched IK for reset (ag_ik_solve; reference Robot.ik_random_restarts agents/robot.py:84-121 via env.py:296): the returned
joint angles are inside the limits and bring the end effector's link frame to the target within the reference's success
threshold, checked with the independent numpy FK of assistive_gym_b200/kinematics.py."""
import numpy as np
import pytest

from assistive_gym_b200 import capi
from assistive_gym_b200.feeding_batch import JACO
from assistive_gym_b200.kinematics import q_from_rpy
from assistive_gym_b200.sim import BatchSim


def _check(fb, sim):
    n = sim.n
    rng = np.random.default_rng(0)
    target = np.array([-0.15, -0.65, 1.15]) + rng.uniform(-0.05, 0.05, size=(n, 3))
    tq = q_from_rpy(JACO['ee_orient_rpy'])
    mask = np.ones(n, dtype=np.int32); mask[1] = 0
    q, err = sim.ik_solve(fb.arm_links, fb.ee_link, target, tq, max_restarts=20, iters=120, threshold=0.01, seed=7, mask=mask)
    ok = mask.astype(bool)
    assert np.all(err[ok] < 0.01), err
    assert np.all(q[ok] >= fb.arm_lower - 1e-6) and np.all(q[ok] <= fb.arm_upper + 1e-6)
    qfull = np.zeros((n, fb.kin.nl)); qfull[:, np.array(JACO['arm']) + 1] = q
    pos, quat = fb.kin.fk(np.broadcast_to(fb.robot_base_pos, (n, 3)), np.broadcast_to(fb.robot_base_quat, (n, 4)), qfull)
    ee = JACO['ee'] + 1
    assert np.linalg.norm(pos[ok, ee] - target[ok], axis=1).max() < 0.0101
    oe = np.minimum(np.linalg.norm(quat[ok, ee] - tq, axis=1), np.linalg.norm(quat[ok, ee] + tq, axis=1))
    assert oe.max() < 0.0101
    # same seed -> same answer; an unreachable target reports its error instead of pretending
    q2, err2 = sim.ik_solve(fb.arm_links, fb.ee_link, target, tq, max_restarts=20, iters=120, threshold=0.01, seed=7, mask=mask)
    assert np.array_equal(q[ok], q2[ok])
    far = np.tile([5.0, 5.0, 5.0], (n, 1))
    _, err3 = sim.ik_solve(fb.arm_links, fb.ee_link, far, tq, max_restarts=2, iters=30, threshold=0.01, seed=1)
    assert np.all(err3 > 1.0)
    with pytest.raises(RuntimeError):
        sim.ik_solve([fb.gl(fb.tool, -1)], fb.ee_link, target, tq)          # a link that is not on the arm's chain


def test_ik_cpu_harness(feeding, emu_lib):
    _check(feeding, BatchSim(feeding.scene, capi.default_config(), 64, _lib=emu_lib))


@pytest.mark.gpu
def test_ik_gpu(feeding, gpu_lib):
    _check(feeding, BatchSim(feeding.scene, capi.default_config(), 4096, device=0))
