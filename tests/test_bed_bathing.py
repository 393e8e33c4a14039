"""SURVEY.md §8(a) row B1 — BedBathing extras on the same backend: wiping targets (util.capsule_points), the
wiper-on-arm contact bookkeeping of `get_total_force`, the 24-float observation; product vs the CPU oracle on the
§8(d) C2 case (wiper pressed into the forearm so that contact rows are live)."""
import numpy as np
import pytest

from assistive_gym_b200 import capi
from assistive_gym_b200.bed_bathing_batch import ARM_DIMS, R_ELBOW, R_WRIST, SAWYER, BedBathingBatch, capsule_points
from assistive_gym_b200.sim import BatchSim
from oracle.oracle_py import OracleSim


@pytest.fixture(scope='module')
def bathing():
    return BedBathingBatch()


def test_capsule_points_counts_and_geometry():
    """bed_bathing.py:176-184 with util.py:80-113: 129 targets on the male arm, 91 on the female;
    every point sits on the cylinder of the capsule, rings 0.03 apart."""
    counts = {}
    for g, (ul, ur, fl, fr) in ARM_DIMS.items():
        u, f = capsule_points([0, 0, 0], [0, 0, -ul], ur, 0.03), capsule_points([0, 0, 0], [0, 0, -fl], fr, 0.03)
        counts[g] = len(u) + len(f)
        assert np.allclose(np.linalg.norm(u[:, :2], axis=1), ur) and np.allclose(np.linalg.norm(f[:, :2], axis=1), fr)
        assert u[:, 2].max() < 0 and u[:, 2].min() > -ul
    assert counts == {'male': 129, 'female': 91}


def test_scene_recipe(bathing):
    bb, sc = bathing, bathing.scene
    assert sc.n_bodies == 6                                   # plane, bed, 2 humans, sawyer, wiper
    assert [j + 1 for j in SAWYER['arm']] == [4, 9, 10, 11, 12, 14, 17]
    assert int(sc['body_nlinks'][bb.robot]) == 25 and int(sc['body_nlinks'][bb.tool]) == 3
    # Sawyer: 7 arm + 2 gripper DoF are live; both humans are fully static in this task
    assert sum(sc['link_jtype'][bb.arm_links] == 1) == 7


def _pressed_pair(bb, make_sim, n, seed):
    """Oracle and product in the same state: reset on the oracle, copy, then drive the wiper onto the forearm."""
    cfg = capi.default_config(residual_threshold=0.0)
    cpu, dev = OracleSim(bb.scene, cfg, n, threads=4), make_sim(bb.scene, cfg, n)
    s = bb.reset(cpu, np.random.default_rng(seed))
    bb.reset(dev, np.random.default_rng(seed), sample=s)
    # start pose: end effector 12 cm above the middle of the forearm, then a motor target 14 cm lower
    male = s['male'].astype(bool)
    mid = np.zeros((n, 3))
    for g, hb in bb.humans.items():
        ls = cpu.get_link_states([bb.gl(hb, R_ELBOW), bb.gl(hb, R_WRIST)])['pos']
        on = male if g == 'male' else ~male
        mid[on] = 0.5 * (ls[on, 0] + ls[on, 1])
    rng = np.random.default_rng(seed + 1)
    arm = np.array(SAWYER['arm']) + 1

    def put(sim, q):
        qfull = q.copy(); qfull[:, np.array(SAWYER['gripper']) + 1] = SAWYER['gripper_pos']
        sim.set_joint_state(bb.arm_links, q=q[:, arm], qd=np.zeros((n, 7)))
        bb.place_tool(sim, bb.base_pos, bb.base_quat, qfull)
        sim.forward_kinematics()

    # hover 25 cm above the forearm, measure the wiper-arm gap there, then start 5 mm above the skin
    q_hi, e_hi = bb.solve_ik(bb.base_pos, bb.base_quat, mid + [0, 0, 0.25], rng, max_restarts=12)
    put(cpu, q_hi)
    gap = np.full(n, np.inf)
    for hb in bb.humans.values():
        c, k = cpu.closest_points(bb.tool, hb, 1.0, max_pts=32)
        gap = np.minimum(gap, np.where(np.arange(32)[None, :] < k[:, None], c['distance'], np.inf).min(axis=1))
    h0 = 0.25 - (gap - 0.005)
    q_hi, e_hi = bb.solve_ik(bb.base_pos, bb.base_quat, mid + np.stack([0 * h0, 0 * h0, h0], axis=1), rng, max_restarts=12)
    # the pressing pose: 5 cm lower, continued from the start pose so that it is the neighbouring IK branch
    from assistive_gym_b200.kinematics import ik_dls, q_from_rpy
    tq = np.broadcast_to(q_from_rpy(SAWYER['ee_orient_rpy']), (n, 4)).copy()
    q_lo, pe, oe = ik_dls(bb.kin, bb.base_pos, bb.base_quat, q_hi.copy(), arm, SAWYER['ee'] + 1, mid + np.stack([0 * h0, 0 * h0, h0 - 0.05], axis=1), tq,
                          bb.arm_lower, bb.arm_upper, iters=200)
    e_lo = np.maximum(pe, oe)
    for sim in (cpu, dev):
        put(sim, q_hi)
        sim.set_motor(bb.arm_links, 1, target=q_lo[:, arm], kp=[0.1] * 7, kd=[1.0] * 7, max_force=[5.0] * 7)
    dev.state_set(cpu.state_get())
    return cpu, dev, s, (e_hi, e_lo, q_lo)


def _check_wiper_on_arm(bb, make_sim, n):
    cpu, dev, s, ik = _pressed_pair(bb, make_sim, n, seed=4)
    tw, alive = bb.targets_world(cpu, s)
    alive_c, alive_d = alive.copy(), alive.copy()
    wiped_c = np.zeros(n, int); wiped_d = np.zeros(n, int)
    fmax = 0.0; rel = 0.0; dq = 0.0
    for i in range(60):
        cpu.step(1); dev.step(1)
        tf_c, th_c, tot_c, new_c = bb.total_force(cpu, tw, alive_c)
        tf_d, th_d, tot_d, new_d = bb.total_force(dev, tw, alive_d)
        wiped_c += new_c; wiped_d += new_d
        big = th_c > 0.5
        if big.any():
            fmax = max(fmax, th_c.max())
            rel = max(rel, (np.abs(th_c - th_d)[big] / th_c[big]).max())
        dq = max(dq, np.abs(cpu.get_joint_states(bb.arm_links)[0] - dev.get_joint_states(bb.arm_links)[0]).max())
    res = dict(force=fmax, force_rel=rel, dq=dq, wiped_cpu=wiped_c.tolist(), wiped_dev=wiped_d.tolist(), ik=np.round(np.maximum(ik[0], ik[1]), 3).tolist())
    print('wiper on arm', res)
    assert fmax > 0.5, res                                      # the case must produce live cloth-on-arm contact
    assert rel < 0.05 and dq < 1e-4, res                        # north-star tolerances: 5 % force, 1e-4 rad
    assert wiped_c.sum() > 0 and np.array_equal(alive_c, alive_d), res      # the same targets are wiped


def _check_env_surface(lib, n):
    from assistive_gym_b200 import envs
    env = envs.make('assistive_gym:BedBathingSawyer-v1', n_envs=n, seed=5)
    env._sim_lib = lib
    obs = env.reset()
    assert env.action_space.shape == (7,) and env.observation_space.shape == (24,)        # bed_bathing.py:10: 17 + 7
    obs = np.atleast_2d(obs)
    assert obs.shape == (n, 24) and np.all(np.isfinite(obs))
    assert set(np.atleast_1d(env.total_target_count).tolist()) <= {129, 91}
    rng = np.random.default_rng(0)
    for _ in range(3):
        a = rng.uniform(-1, 1, size=(n, 7))
        o, r, d, info = env.step(a if n > 1 else a[0])
        assert np.all(np.isfinite(np.atleast_2d(o))) and np.all(np.isfinite(np.atleast_1d(r)))
    assert set(info) >= {'total_force_on_human', 'task_success', 'action_robot_len', 'obs_robot_len'}
    env.close()


def _check_fused_vs_api(lib, n):
    """Fused kernels vs the per-call API path from identical resets; a strong arm motor makes the wiper touch the arm."""
    from assistive_gym_b200 import envs
    a, b = (envs.make('assistive_gym:BedBathingSawyer-v1', n_envs=n, seed=9) for _ in range(2))
    a._sim_lib = b._sim_lib = lib
    oa, ob = np.atleast_2d(a.reset()), np.atleast_2d(b.reset())
    assert np.allclose(oa, ob, atol=1e-6)
    rng = np.random.default_rng(2)
    for k in range(5):
        act = rng.uniform(-1, 1, size=(n, 7)).astype(np.float32)
        o1, r1, d1, i1 = a.step(act if n > 1 else act[0])
        o2, r2, d2, i2 = b.step_reference_api(act)
        # the two paths round the PD targets differently (fp32 on the device, float64 on the host); envs whose arm or
        # wiper rests against the bed amplify that, so the tight bound is asserted on the median env
        eo = np.abs(np.atleast_2d(o1) - np.atleast_2d(o2)).max(axis=1)
        er = np.abs(np.atleast_1d(r1) - np.atleast_1d(r2))
        assert np.median(eo) < 1e-4 and eo.max() < 5e-3, (k, eo)
        assert np.median(er) < 1e-3 and er.max() < 5e-2, (k, er)
        assert np.array_equal(np.atleast_1d(i1['task_success']), np.atleast_1d(i2['task_success']))
    a.close(); b.close()


def _check_fused_wiping(bb, make_sim, n):
    """The pressed-wiper case through the fused kernels (device) against `total_force` on the oracle."""
    cpu, dev, s, ik = _pressed_pair(bb, make_sim, n, seed=4)
    tw, alive = bb.targets_world(cpu, s)
    alive_c = alive.copy()
    dev.bathing_init(bb.bathing_params(), s['male'], tw, alive)
    from tests.parity_cases import take_step_targets
    arm = np.array(SAWYER['arm']) + 1
    q_lo = ik[2][:, arm]
    wiped_c = np.zeros(n, int); wiped_d = np.zeros(n, int)
    for k in range(12):
        q = cpu.get_joint_states(bb.arm_links)[0]
        act = np.clip((q_lo - q) / 0.25, -1, 1).astype(np.float32)          # drive towards the pressing pose
        cpu.set_motor_targets(bb.arm_links, take_step_targets(q, act, bb.arm_lower, bb.arm_upper))
        cpu.step(5)
        obs, rew, done, info = dev.bathing_step_host(act)
        tf, th, tot, new = bb.total_force(cpu, tw, alive_c)
        wiped_c += new; wiped_d += info[:, 3].astype(int)
        assert np.abs(info[:, 0] - tot).max() < 0.05 * max(1.0, tot.max()), (k, info[:, 0], tot)
        assert np.abs(obs[:, 7:14] - ((q_now := cpu.get_joint_states(bb.arm_links)[0]) + np.pi) % (2 * np.pi) + np.pi).max() < 1e-4
    print('fused wiping', wiped_c.tolist(), wiped_d.tolist())
    assert wiped_c.sum() > 0 and np.array_equal(wiped_c, wiped_d)


def test_wiper_on_arm_cpu_harness(bathing, emu_lib):
    _check_wiper_on_arm(bathing, lambda sc, cfg, n: BatchSim(sc, cfg, n, _lib=emu_lib), 2)


def test_env_surface_cpu_harness(emu_lib):
    _check_env_surface(emu_lib, 3)


def test_fused_vs_api_cpu_harness(emu_lib):
    _check_fused_vs_api(emu_lib, 5)


def test_fused_wiping_cpu_harness(bathing, emu_lib):
    _check_fused_wiping(bathing, lambda sc, cfg, n: BatchSim(sc, cfg, n, _lib=emu_lib), 2)


@pytest.mark.gpu
def test_wiper_on_arm_gpu(bathing, gpu_lib):
    _check_wiper_on_arm(bathing, lambda sc, cfg, n: BatchSim(sc, cfg, n, device=0), 8)


@pytest.mark.gpu
def test_env_surface_gpu(gpu_lib):
    _check_env_surface(None, 16)


@pytest.mark.gpu
def test_fused_vs_api_gpu(gpu_lib):
    _check_fused_vs_api(None, 16)


@pytest.mark.gpu
def test_fused_wiping_gpu(bathing, gpu_lib):
    _check_fused_wiping(bathing, lambda sc, cfg, n: BatchSim(sc, cfg, n, device=0), 8)
