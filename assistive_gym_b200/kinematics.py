"""Batched host-side kinematics (numpy) used at reset time: forward kinematics over a scene body
and damped-least-squares inverse kinematics with random restarts.

Reset-time helpers only — the reference does the same work through `p.getLinkState`,
`p.calculateInverseKinematics` (agents/agent.py:252-273) and `Robot.ik_random_restarts`
(agents/robot.py:84-121).  SURVEY.md §8(f) lists batched reset as the first "next" row; the
stepping hot path never calls into this module.
"""
import numpy as np


def q_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def q_conj(q):
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def _cross(a, b):
    """np.cross for [..., 3] operands without its axis shuffling (the IK loop calls this ~10^4 times per reset)."""
    a, b = np.broadcast_arrays(a, b)
    out = np.empty(a.shape, dtype=np.result_type(a, b))
    out[..., 0] = a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1]
    out[..., 1] = a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2]
    out[..., 2] = a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]
    return out


def q_rot(q, v):
    u = q[..., :3]
    t = 2.0 * _cross(u, v)
    return v + q[..., 3:4] * t + _cross(u, t)


def q_axis(axis, ang):
    s = np.sin(0.5 * ang)[..., None]
    return np.concatenate([axis * s, np.cos(0.5 * ang)[..., None]], axis=-1)


def q_from_rpy(rpy):
    rpy = np.asarray(rpy, dtype=np.float64)
    r, p, y = rpy[..., 0], rpy[..., 1], rpy[..., 2]
    cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(p / 2), np.sin(p / 2), np.cos(y / 2), np.sin(y / 2)
    return np.stack([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                     cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy], axis=-1)


class BodyKinematics:
    """Forward kinematics / geometric Jacobian of one body of a finalized scene, batched over envs."""

    def __init__(self, scene, body):
        self.l0 = int(scene['body_link0'][body])
        self.nl = int(scene['body_nlinks'][body])
        sl = slice(self.l0, self.l0 + self.nl)
        self.parent = scene['link_parent'][sl] - self.l0
        self.jtype = scene['link_jtype'][sl]
        self.axis = scene['link_axis'][sl]
        self.jpos = scene['link_jpos'][sl]
        self.jquat = scene['link_jquat'][sl]
        self.com = scene['link_com'][sl]
        self.iquat = scene['link_iquat'][sl]
        self.lower = scene['link_lower'][sl]
        self.upper = scene['link_upper'][sl]

    def fk(self, base_pos, base_quat, q, upto=None):
        """q: [N, nl] joint values indexed by local link (column 0 unused).  Returns pos [N,nl,3], quat [N,nl,4].
        `upto`: only links 0..upto are computed (links are in DFS pre-order, so a link's ancestors precede it)."""
        N = q.shape[0]
        pos = np.zeros((N, self.nl, 3))
        quat = np.zeros((N, self.nl, 4))
        pos[:, 0] = base_pos
        quat[:, 0] = base_quat
        for k in range(1, self.nl if upto is None else upto + 1):
            p = self.parent[k]
            jp = pos[:, p] + q_rot(quat[:, p], self.jpos[k])
            jq = q_mul(quat[:, p], np.broadcast_to(self.jquat[k], (N, 4)))
            if self.jtype[k] == 1:
                jq = q_mul(jq, q_axis(self.axis[k], q[:, k]))
            elif self.jtype[k] == 2:
                jp = jp + q_rot(jq, self.axis[k] * q[:, k, None])
            pos[:, k] = jp
            quat[:, k] = jq / np.linalg.norm(jq, axis=-1, keepdims=True)
        return pos, quat

    def link_com_pose(self, pos, quat, k):
        return pos[:, k] + q_rot(quat[:, k], self.com[k]), q_mul(quat[:, k], np.broadcast_to(self.iquat[k], quat[:, k].shape))

    def jacobian(self, pos, quat, k, point, joints):
        """Geometric Jacobian [N, 6, len(joints)] of `point` (world, [N,3]) on link k w.r.t. `joints` (local link ids)."""
        N = pos.shape[0]
        J = np.zeros((N, 6, len(joints)))
        anc = set()
        j = k
        while j > 0:
            anc.add(j)
            j = self.parent[j]
        for c, jl in enumerate(joints):
            if jl not in anc:
                continue
            a = q_rot(quat[:, jl], self.axis[jl])
            if self.jtype[jl] == 1:
                J[:, :3, c] = _cross(a, point - pos[:, jl])
                J[:, 3:, c] = a
            elif self.jtype[jl] == 2:
                J[:, :3, c] = a
        return J


def ik_dls(kin, base_pos, base_quat, q_init, joints, ee, target_pos, target_quat, lower, upper,
           iters=200, damping=0.05, step_clip=0.2, tol=1e-5):
    """Damped least squares IK for link `ee` (link frame).  q_init [N, nl]; returns q [N, nl], pos_err, ori_err.
    Envs whose 6-D error has dropped below `tol` leave the active set, so late iterations only touch stragglers."""
    q = q_init.copy()
    lam2I = damping ** 2 * np.eye(6)
    base_pos, base_quat = np.broadcast_to(base_pos, (q.shape[0], 3)), np.broadcast_to(base_quat, (q.shape[0], 4))
    act = np.arange(q.shape[0])
    for _ in range(iters):
        qa = q[act]
        pos, quat = kin.fk(base_pos[act], base_quat[act], qa, upto=ee)
        ep = target_pos[act] - pos[:, ee]
        qe = q_mul(target_quat[act], q_conj(quat[:, ee]))
        qe = qe * np.where(qe[:, 3:4] < 0, -1.0, 1.0)
        err = np.concatenate([ep, 2.0 * qe[:, :3]], axis=1)
        live = np.abs(err).max(axis=1) > tol
        if not live.any():
            break
        if not live.all():
            act, qa, pos, quat, err = act[live], qa[live], pos[live], quat[live], err[live]
        J = kin.jacobian(pos, quat, ee, pos[:, ee], joints)
        Jt = np.transpose(J, (0, 2, 1))
        dq = np.einsum('nij,nj->ni', Jt, np.linalg.solve(J @ Jt + lam2I, err[..., None])[..., 0])
        qa[:, joints] = np.clip(qa[:, joints] + np.clip(dq, -step_clip, step_clip), lower, upper)
        q[act] = qa
    pos, quat = kin.fk(base_pos, base_quat, q, upto=ee)
    pe = np.linalg.norm(target_pos - pos[:, ee], axis=1)
    oe = np.minimum(np.linalg.norm(target_quat - quat[:, ee], axis=1), np.linalg.norm(target_quat + quat[:, ee], axis=1))
    return q, pe, oe
