"""Dressing parity cases: the fused DressingPR2 step of the product next to a numpy restatement of reference
envs/dressing.py:12-106 + envs/util.py:125-202 + envs/env.py:174-274 driven through the CPU oracle's per-call API."""
import numpy as np

from assistive_gym_b200.dressing_batch import CLOTH_ANCHORS, L_ELBOW, L_SHOULDER, L_WRIST, RADII, TRIANGLE1, TRIANGLE2, DressingBatch
from assistive_gym_b200.kinematics import q_conj, q_mul, q_rot


def _sign(v):
    return np.sign(v)


def _signed_volume(a, b, c, d):
    return (1.0 / 6.0) * np.dot(np.cross(b - a, c - a), d - a)


def line_intersects_triangle(p0, p1, p2, q0, q1):                      # util.py:125-132
    if _sign(_signed_volume(q0, p0, p1, p2)) != _sign(_signed_volume(q1, p0, p1, p2)):
        return _sign(_signed_volume(q0, q1, p0, p1)) == _sign(_signed_volume(q0, q1, p1, p2)) == _sign(_signed_volume(q0, q1, p2, p0))
    return False


def sleeve_on_arm_reward(t1, t2, shoulder, elbow, wrist, hand_r, elbow_r, shoulder_r):      # util.py:134-202
    hand_end = wrist + (wrist - elbow) / np.linalg.norm(wrist - elbow) * hand_r * 2
    elbow_end = elbow + (elbow - wrist) / np.linalg.norm(wrist - elbow) * elbow_r
    shoulder_end = shoulder + (shoulder - elbow) / np.linalg.norm(shoulder - elbow) * shoulder_r
    pts = np.concatenate([t1, t2], axis=0)

    def around(normal, origin):
        normal = normal / np.linalg.norm(normal)
        t = np.cross(np.array([1, 1, 0]), normal); t = t / np.linalg.norm(t)
        b = np.cross(t, normal); b = b / np.linalg.norm(b)
        tp, bp = (pts - origin) @ t, (pts - origin) @ b
        return np.any(tp > 0) and np.any(tp < 0) and np.any(bp > 0) and np.any(bp < 0)
    fa, ua = around(hand_end - elbow_end, hand_end), around(elbow_end - shoulder_end, shoulder_end)
    fh = line_intersects_triangle(*t1, hand_end, elbow_end) or line_intersects_triangle(*t2, hand_end, elbow_end)
    uh = line_intersects_triangle(*t1, elbow_end, shoulder_end) or line_intersects_triangle(*t2, elbow_end, shoulder_end)
    centre = pts.mean(axis=0)
    return (fa and fh, ua and uh, np.linalg.norm(centre - hand_end), np.linalg.norm(centre - elbow), np.linalg.norm(hand_end - centre),
            np.linalg.norm(hand_end - elbow_end), np.linalg.norm(elbow - shoulder))


class DressingReference:
    """DressingEnv.step (dressing.py:12-77) on a per-call simulation API (the CPU oracle), batched over envs."""

    def __init__(self, db, sim, male, sample=None):
        self.db, self.sim, self.male = db, sim, np.asarray(male).astype(bool)
        self.iteration = 0
        self.task_success = np.zeros(sim.n)
        imp = None if sample is None else sample.get('impairment')
        self.tremor_on = np.zeros(sim.n, dtype=bool) if imp is None else (imp == 3)
        self.tremors = np.zeros((sim.n, 10)) if sample is None or 'tremors' not in sample else sample['tremors']

    def step(self, action):
        db, sim, n = self.db, self.sim, self.sim.n
        self.iteration += 1
        a = np.clip(np.asarray(action, dtype=np.float64), -1, 1) * 0.05
        q = sim.get_joint_states(db.arm_links)[0].astype(np.float64)
        act = a.copy()
        for _ in range(5):                                                # env.py:202-217
            below, above = q + act < db.arm_lower, q + act > db.arm_upper
            act[below] = 0; act[above] = 0
            q = np.where(below, db.arm_lower, q); q = np.where(above, db.arm_upper, q)
            q = q + act
        sim.set_motor_targets(db.arm_links, q)
        if self.tremor_on.any():                                          # env.py:212-215: the tremor human's targets flip every env step
            sgn = 1.0 if self.iteration % 2 == 0 else -1.0
            for g, hb in db.humans.items():
                sel = self.tremor_on & (self.male if g == 'male' else ~self.male)
                al = db.human_arm_links[g]
                cur = db.human_rest.copy()
                tgt = np.where(sel[:, None], db.human_rest + sgn * self.tremors, cur)
                if sel.any():
                    # envs of the other gender / without tremor keep their rest targets (set at reset)
                    sim.set_motor_targets(al, tgt)
        for _ in range(5):                                                # env.py:223-231 + dressing.py:200-210
            sim.step(1)
            sim.cloth_anchor_follow(db.ee_link)
        x, _ = sim.cloth_get_state()
        cnt, node, cpos, force, link = sim.cloth_get_contacts(2048)
        ee = sim.get_link_states([db.ee_link])
        ep, eq = ee['pos'][:, 0].astype(np.float64), ee['quat'][:, 0].astype(np.float64)
        rb = int(db.scene['body_link0'][db.robot])
        rs = sim.get_link_states([rb])
        rp, rq = rs['com_pos'][:, 0].astype(np.float64), rs['com_quat'][:, 0].astype(np.float64)
        rqi = q_conj(rq)
        obs = np.zeros((n, 24)); rew = np.zeros(n); info = np.zeros((n, 4))
        qa = sim.get_joint_states(db.arm_links)[0].astype(np.float64)
        obs[:, 0:3] = q_rot(rqi, ep - rp); obs[:, 3:7] = q_mul(rqi, eq)
        obs[:, 7:14] = (qa + np.pi) % (2 * np.pi) - np.pi
        limb = np.zeros((n, 3, 3))
        for g, hb in db.humans.items():
            sel = self.male if g == 'male' else ~self.male
            ls = sim.get_link_states([db.gl(hb, L_SHOULDER), db.gl(hb, L_ELBOW), db.gl(hb, L_WRIST)])['pos']
            limb[sel] = ls[sel]
        for j in range(3):
            obs[:, 14 + 3 * j:17 + 3 * j] = q_rot(rqi, limb[:, j] - rp)
        vel = np.linalg.norm(ee['lin_vel'][:, 0], axis=1)
        robot_on_human = sum(sim.contact_force_sum(db.robot, hb) for hb in db.humans.values())
        for e in range(n):
            hr, er, sr = RADII['male' if self.male[e] else 'female']
            fin, uin, d_fore, d_upper, d_hand, fore_len, upper_len = sleeve_on_arm_reward(x[e, TRIANGLE1], x[e, TRIANGLE2], limb[e, 0], limb[e, 1], limb[e, 2], hr, er, sr)
            if uin:
                rd = fore_len + (d_upper if d_upper < upper_len else 0.0)
            elif fin and d_fore < fore_len:
                rd = d_fore
            else:
                rd = -d_hand
            f = np.linalg.norm(force[e, :cnt[e]] * 10.0, axis=1)
            keep = (cpos[e, :cnt[e], 2] < ep[e, 2] - 0.05) & (f < 20)
            cs = f[keep].sum()
            obs[e, 23] = cs
            pref = 0.25 * (-vel[e]) + 0.01 * (-cs)
            rew[e] = 1.0 * rd + 0.01 * (-np.linalg.norm(action[e])) + pref
            self.task_success[e] = max(self.task_success[e], rd)
            info[e] = [robot_on_human[e] + cs, float(self.task_success[e] >= 0.4), rd, (1 if fin else 0) + (2 if uin else 0)]
        return obs, rew, np.full(n, float(self.iteration >= 200)), info
