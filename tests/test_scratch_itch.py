"""ScratchItchJaco-v1: the fused step of the product against a numpy restatement of reference
envs/scratch_itch.py:10-91 driven through the CPU oracle's per-call API, from the same reset."""
import numpy as np
import pytest

from assistive_gym_b200 import capi
from assistive_gym_b200.kinematics import q_conj, q_mul, q_rot
from assistive_gym_b200.scratch_itch_batch import R_ELBOW, R_SHOULDER, R_WRIST, ScratchItchBatch
from assistive_gym_b200.sim import BatchSim
from oracle.oracle_py import OracleSim


class ScratchReference:
    def __init__(self, sb, sim, smp):
        self.sb, self.sim, self.s = sb, sim, smp
        self.male = smp['male'].astype(bool)
        self.iteration = 0
        self.task_success = np.zeros(sim.n, dtype=int)
        self.prev = np.zeros((sim.n, 3))
        self.limb = sb.limb_links(smp)

    def step(self, action):
        sb, sim, n = self.sb, self.sim, self.sim.n
        self.iteration += 1
        a = np.clip(np.asarray(action, dtype=np.float64), -1, 1) * 0.05
        q = sim.get_joint_states(sb.arm_links)[0].astype(np.float64)
        act = a.copy()
        for _ in range(5):
            below, above = q + act < sb.arm_lower, q + act > sb.arm_upper
            act[below] = 0; act[above] = 0
            q = np.where(below, sb.arm_lower, q); q = np.where(above, sb.arm_upper, q)
            q = q + act
        sim.set_motor_targets(sb.arm_links, q)
        sim.step(5)
        idx = np.arange(n)
        ll = sim.get_link_states(list(self.limb))
        target = ll['pos'][idx, idx].astype(np.float64) + q_rot(ll['quat'][idx, idx].astype(np.float64), self.s['target_local'])
        tip = sb.gl(sb.tool, 1)
        ts = sim.get_link_states([tip, sb.ee_link])
        tp, tq = ts['pos'][:, 0].astype(np.float64), ts['quat'][:, 0].astype(np.float64)
        rs = sim.get_link_states([int(sb.scene['body_link0'][sb.robot])])
        rp, rqi = rs['com_pos'][:, 0].astype(np.float64), q_conj(rs['com_quat'][:, 0].astype(np.float64))
        obs = np.zeros((n, 30)); rew = np.zeros(n); info = np.zeros((n, 4))
        tp_r, tg_r = q_rot(rqi, tp - rp), q_rot(rqi, target - rp)
        obs[:, 0:3], obs[:, 3:7], obs[:, 7:10], obs[:, 10:13] = tp_r, q_mul(rqi, tq), tp_r - tg_r, tg_r
        qa = sim.get_joint_states(sb.arm_links)[0].astype(np.float64)
        obs[:, 13:20] = (qa + np.pi) % (2 * np.pi) - np.pi
        for g, hb in sb.humans.items():
            sel = self.male if g == 'male' else ~self.male
            ls = sim.get_link_states([sb.gl(hb, R_SHOULDER), sb.gl(hb, R_ELBOW), sb.gl(hb, R_WRIST)])['pos']
            for j in range(3):
                obs[sel, 20 + 3 * j:23 + 3 * j] = q_rot(rqi[sel], ls[sel, j] - rp[sel])
        tool_force = sim.contact_force_sum(sb.tool).astype(np.float64)
        obs[:, 29] = tool_force
        vel = np.linalg.norm(ts['lin_vel'][:, 1], axis=1)
        for e in range(n):
            hb = sb.humans['male' if self.male[e] else 'female']
            total = float(sim.contact_force_sum(sb.robot, hb)[e])
            c, k = sim.get_contacts(sb.tool, hb, max_pts=32)
            at_target, cpos = 0.0, None
            for i in range(k[e]):
                f = float(c['normal_force'][e, i])
                total += f
                if int(c['link_a'][e, i]) in (sb.gl(sb.tool, 0), tip) and np.linalg.norm(c['pos_b'][e, i] - target[e]) < 0.025:
                    at_target += f; cpos = c['pos_b'][e, i].astype(np.float64)
            scratch = 0.0
            if cpos is not None and np.linalg.norm(cpos - self.prev[e]) > 0.01 and at_target < 10:
                scratch = 5.0; self.prev[e] = cpos; self.task_success[e] += 1
            pref = 0.25 * (-vel[e]) + 0.01 * (-(total - at_target)) + 0.05 * (0.0 if at_target < 10 else -at_target)
            rew[e] = -np.linalg.norm(target[e] - tp[e]) + 0.01 * (-np.linalg.norm(action[e])) + scratch + pref
            info[e] = [total, float(self.task_success[e] >= 25), at_target, self.task_success[e]]
        return obs, rew, np.full(n, float(self.iteration >= 200)), info


def _press_tip_on_target(sb, prod, smp, depth=0.004):
    """Move the arm so that the tool tip (a 1 cm sphere) sits `depth` inside the skin at the target: IK of the end effector with its
    CURRENT orientation (the tip's offset from the end effector is then a constant world vector).  Returns the envs where it worked."""
    n = prod.n
    idx = np.arange(n)
    limb = sb.limb_links(smp)
    ll = prod.get_link_states(list(limb))
    lp, lq = ll['pos'][idx, idx].astype(np.float64), ll['quat'][idx, idx].astype(np.float64)
    radial = smp['target_local'] * np.array([1.0, 1.0, 0.0])
    nrm = q_rot(lq, radial / np.linalg.norm(radial, axis=1, keepdims=True))
    target = lp + q_rot(lq, smp['target_local'])
    ts = prod.get_link_states([sb.gl(sb.tool, 1), sb.ee_link])
    tip, ee, eq = ts['pos'][:, 0].astype(np.float64), ts['pos'][:, 1].astype(np.float64), ts['quat'][:, 1].astype(np.float64)
    goal_ee = target + nrm * (0.01 - depth) - (tip - ee)
    q7, err = prod.ik_solve(sb.arm_links, sb.ee_link, goal_ee, eq, max_restarts=30, iters=150, threshold=2e-3, seed=5)
    return q7.astype(np.float64), err < 5e-3


def _contact_case(lib, n=6, steps=3):
    """The tool tip pressed onto the target: tool force at the target, the scratch count and the reward agree with the restatement."""
    sb = ScratchItchBatch()
    cfg = capi.default_config(residual_threshold=0.0)
    prod = BatchSim(sb.scene, cfg, n, _lib=lib)
    smp = sb.reset(prod, np.random.default_rng(3))
    q7, ok = _press_tip_on_target(sb, prod, smp)
    assert ok.sum() >= 2, ok
    smp['q7'] = np.where(ok[:, None], q7, smp['q7'])
    orc = OracleSim(sb.scene, cfg, n, threads=4)
    for s in (prod, orc):
        sb.reset(s, np.random.default_rng(3), sample=smp)
    prod.state_set(orc.state_get().astype(np.float32))
    sb.start_fused(prod, smp)
    ref = ScratchReference(sb, orc, smp)
    seen = 0
    for it in range(steps):
        a = np.zeros((n, 7))
        obs, rew, done, info = prod.scratch_step_host(a.astype(np.float32))
        obs_r, rew_r, done_r, info_r = ref.step(a)
        assert np.array_equal(info[:, 3], info_r[:, 3]), (info[:, 3], info_r[:, 3])                  # scratches counted
        assert np.all(np.abs(info[:, 2] - info_r[:, 2]) <= 0.05 * np.abs(info_r[:, 2]) + 1e-2)       # tool force at the target (5 %)
        assert np.all(np.abs(info[:, 0] - info_r[:, 0]) <= 0.05 * np.abs(info_r[:, 0]) + 1e-2)       # total force on the person
        assert np.all(np.abs(rew - rew_r) < 2e-2 + 0.02 * np.abs(rew_r))          # the reward carries the (5 %) forces
        seen = max(seen, int(info_r[:, 3].max()))
        prod.state_set(orc.state_get().astype(np.float32))
    assert seen >= 1                                                        # the first touch counts (prev_target_contact_pos starts at 0)


def _case(lib, n=4, steps=3):
    sb = ScratchItchBatch()
    cfg = capi.default_config(residual_threshold=0.0)
    prod = BatchSim(sb.scene, cfg, n, _lib=lib)
    smp = sb.reset(prod, np.random.default_rng(2))
    assert sb.unresolved == 0 and sb.ik_err.max() < 0.05          # ik_random_restarts keeps the best restart when none reaches 0.01 (robot.py:113-117)
    orc = OracleSim(sb.scene, cfg, n, threads=4)
    sb.reset(orc, np.random.default_rng(2), sample=smp)
    # press the tool tip onto the target so that the contact / scratch logic is live: aim the arm at the target by IK
    idx = np.arange(n)
    ll = prod.get_link_states(list(sb.limb_links(smp)))
    target = ll['pos'][idx, idx] + q_rot(ll['quat'][idx, idx].astype(np.float64), smp['target_local'])
    prod.state_set(orc.state_get().astype(np.float32))
    sb.start_fused(prod, smp)
    ref = ScratchReference(sb, orc, smp)
    rng = np.random.default_rng(9)
    for it in range(steps):
        a = rng.uniform(-1, 1, size=(n, 7))
        obs, rew, done, info = prod.scratch_step_host(a.astype(np.float32))
        obs_r, rew_r, done_r, info_r = ref.step(a)
        assert np.abs(obs[:, :13] - obs_r[:, :13]).max() < 1e-3              # tool pose / target in the robot frame (1e-3 m)
        assert np.abs(obs[:, 13:20] - obs_r[:, 13:20]).max() < 1e-4          # joint angles (1e-4 rad)
        assert np.abs(obs[:, 20:29] - obs_r[:, 20:29]).max() < 1e-3
        assert np.all(np.abs(obs[:, 29] - obs_r[:, 29]) <= 0.05 * np.abs(obs_r[:, 29]) + 1e-3)      # tool force (5 %)
        assert np.abs(rew - rew_r).max() < 5e-3 and np.array_equal(done, done_r)
        assert np.array_equal(info[:, 3], info_r[:, 3])
        prod.state_set(orc.state_get().astype(np.float32))
    return target


def test_fused_scratch_itch_step_host_compiled(emu_lib):
    _case(emu_lib)


def test_scratch_contact_host_compiled(emu_lib):
    _contact_case(emu_lib)


@pytest.mark.gpu
def test_fused_scratch_itch_step_cuda(gpu_lib):
    _case(gpu_lib, n=8, steps=4)
    _contact_case(gpu_lib, n=8)


def test_scratch_bookkeeping_counts_a_moving_contact(emu_lib):
    """The scratch reward (scratch_itch.py:26-30) with a synthetic state: the tool tip is teleported onto the target, then 2 cm
    along the limb: the first touch and the moved touch each count once, a touch that has not moved does not."""
    from assistive_gym_b200 import envs
    env = envs.make('ScratchItchJaco-v1', n_envs=2)
    env._sim_lib = emu_lib
    env.reset()
    sim, sb = env.id, env._sb
    env.update_targets()
    assert np.all(np.isfinite(env.target_pos)) and env.target_pos.shape == (2, 3)
    limb = sim.get_link_states(list(env._limb_links))
    idx = np.arange(2)
    d = np.linalg.norm(env.target_pos - limb['pos'][idx, idx], axis=1)
    assert np.all(d > 0.02) and np.all(d < 0.3)                                 # on the limb's surface, inside its length
    env.close()
