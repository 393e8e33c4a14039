"""`Agent` — the reference's thin wrapper around one simulated body (envs/agents/agent.py), re-hosted
on the batched C-ABI backend.  Same method names and argument meaning; every query is answered for all
N lock-step envs at once and squeezed to the reference's shapes when N == 1."""
import numpy as np

from ... import scene as sc_util

MOTOR_POSITION, MOTOR_VELOCITY = 1, 2


class Agent:
    def __init__(self):
        self.base = -1
        self.body = None
        self.lower_limits = None
        self.upper_limits = None
        self.ik_lower_limits = None
        self.ik_upper_limits = None
        self.controllable_joint_indices = []

    # ---- wiring
    def init(self, body, sim, np_random, indices=None):
        """`sim` replaces the reference's `physicsClientId` (agent.py:17)."""
        self.body = body
        self.sim = sim
        self.id = sim
        self.np_random = np_random
        s = sim.scene
        self.link0 = int(s['body_link0'][body])
        self.num_joints = int(s['body_nlinks'][body]) - 1
        self.all_joint_indices = list(range(self.num_joints))
        if indices != -1:
            self.update_joint_limits()
            self.controllable_joint_lower_limits = np.array([self.lower_limits[i] for i in self.controllable_joint_indices])
            self.controllable_joint_upper_limits = np.array([self.upper_limits[i] for i in self.controllable_joint_indices])

    def _gl(self, link):
        return self.link0 + 1 + int(link)

    def _sq(self, a):
        return a[0] if self.sim.n == 1 else a

    # ---- control / state (agent.py:28-98)
    def control(self, indices, target_angles, gains, forces):
        n = len(indices)
        gains = [gains] * n if np.isscalar(gains) else list(gains)
        forces = [forces] * n if np.isscalar(forces) else list(forces)
        tgt = np.broadcast_to(np.asarray(target_angles, dtype=np.float32), (self.sim.n, n))
        self.sim.set_motor([self._gl(j) for j in indices], MOTOR_POSITION, target=tgt, kp=gains, kd=[1.0] * n, max_force=forces)

    def get_joint_angles(self, indices=None):
        if indices is None:
            indices = self.all_joint_indices
        elif not len(indices):
            return []
        return self._sq(self.sim.get_joint_states([self._gl(j) for j in indices])[0].astype(np.float64))

    def get_joint_angles_dict(self, indices=None):
        return {j: a for j, a in zip(indices, np.atleast_2d(self.get_joint_angles(indices)).T)}

    def get_pos_orient(self, link, center_of_mass=False, convert_to_realworld=False):
        st = self.sim.get_link_states([self.link0 if link == self.base else self._gl(link)])
        if link == self.base or center_of_mass:
            pos, orient = st['com_pos'][:, 0], st['com_quat'][:, 0]     # PyBullet base pose = inertial frame
        else:
            pos, orient = st['pos'][:, 0], st['quat'][:, 0]
        if convert_to_realworld:
            return self.convert_to_realworld(self._sq(pos), self._sq(orient))
        return self._sq(pos.astype(np.float64)), self._sq(orient.astype(np.float64))

    def convert_to_realworld(self, pos, orient=(0, 0, 0, 1)):
        from ...kinematics import q_conj, q_mul, q_rot
        bp, bq = self.get_base_pos_orient()
        bp, bq = np.atleast_2d(bp), np.atleast_2d(bq)
        pos = np.atleast_2d(np.asarray(pos, dtype=np.float64))
        orient = np.asarray(orient, dtype=np.float64)
        if orient.shape[-1] == 3:
            orient = self.get_quaternion(orient)
        orient = np.broadcast_to(np.atleast_2d(orient), (pos.shape[0], 4))
        qi = q_conj(bq)
        return self._sq(q_rot(qi, pos - bp)), self._sq(q_mul(qi, orient))

    def get_base_pos_orient(self):
        return self.get_pos_orient(self.base)

    def get_velocity(self, link):
        st = self.sim.get_link_states([self.link0 if link == self.base else self._gl(link)])
        return self._sq(st['lin_vel'][:, 0].astype(np.float64))

    def get_euler(self, quaternion):
        return sc_util.euler_from_quat(np.asarray(quaternion, dtype=np.float64))

    def get_quaternion(self, euler):
        return sc_util.quat_from_rpy(np.asarray(euler, dtype=np.float64))

    def get_mass(self, link):
        return float(self.sim.scene['link_mass'][self.link0 if link == self.base else self._gl(link)])

    def get_motor_joint_states(self, joints=None):
        joints = self.all_joint_indices if joints is None else joints
        jt = self.sim.scene['link_jtype']
        motor = [j for j in joints if jt[self._gl(j)] in (1, 2)]
        q, qd, tau = self.sim.get_joint_states([self._gl(j) for j in motor])
        return motor, self._sq(q), self._sq(qd), self._sq(tau)

    # ---- contacts (agent.py:100-130)
    def get_contact_points(self, agentB=None, linkA=None, linkB=None, max_pts=64):
        c, n = self.sim.get_contacts(self.body, -2 if agentB is None else agentB.body,
                                     -2 if linkA is None else linkA, -2 if linkB is None else linkB, max_pts=max_pts)
        out = []
        for e in range(self.sim.n):
            r = c[e, :n[e]]
            la = [int(x) - self.link0 - 1 for x in r['link_a']]
            lb = [int(x) - int(self.sim.scene['body_link0'][self.sim.scene['link_body'][int(x)]]) - 1 for x in r['link_b']]
            out.append((la, lb, [p for p in r['pos_a']], [p for p in r['pos_b']], [float(f) for f in r['normal_force']]))
        return out[0] if self.sim.n == 1 else out

    def get_closest_points(self, agentB, distance=4.0, linkA=None, linkB=None, max_pts=64):
        c, n = self.sim.closest_points(self.body, agentB.body, distance, max_pts=max_pts)
        out = []
        for e in range(self.sim.n):
            r = c[e, :min(n[e], max_pts)]
            la = [int(x) - self.link0 - 1 for x in r['link_a']]
            lb = [int(x) - agentB.link0 - 1 for x in r['link_b']]
            keep = [i for i in range(len(r)) if (linkA is None or la[i] == linkA) and (linkB is None or lb[i] == linkB)]
            out.append(([la[i] for i in keep], [lb[i] for i in keep], [r['pos_a'][i] for i in keep], [r['pos_b'][i] for i in keep],
                        [float(r['distance'][i]) for i in keep]))
        return out[0] if self.sim.n == 1 else out

    # ---- setters (agent.py:145-200)
    def set_base_pos_orient(self, pos, orient):
        orient = np.asarray(orient, dtype=np.float64)
        if orient.shape[-1] == 3:
            orient = self.get_quaternion(orient)
        # PyBullet positions the inertial frame; the backend stores the link frame
        from ...kinematics import q_rot
        com = self.sim.scene['link_com'][self.link0]
        pos = np.asarray(pos, dtype=np.float64) - q_rot(np.atleast_2d(orient), com)
        self.sim.set_base_pose(self.body, pos, orient)
        self.sim.forward_kinematics()

    def set_base_velocity(self, linear_velocity, angular_velocity):
        self.sim.set_base_velocity(self.body, linear_velocity, angular_velocity)

    def set_joint_angles(self, indices, angles, use_limits=True, velocities=0):
        angles = np.broadcast_to(np.asarray(angles, dtype=np.float64), (self.sim.n, len(indices))).copy()
        if use_limits:
            lo = np.array([self.lower_limits[j] for j in indices])
            hi = np.array([self.upper_limits[j] for j in indices])
            angles = np.clip(angles, lo, hi)
        vel = np.broadcast_to(np.asarray(velocities, dtype=np.float64), angles.shape)
        self.sim.set_joint_state([self._gl(j) for j in indices], q=angles, qd=vel)
        self.sim.forward_kinematics()

    def reset_joints(self):
        self.set_joint_angles(self.all_joint_indices, [0] * len(self.all_joint_indices))

    def set_frictions(self, links, lateral_friction=None, spinning_friction=None, rolling_friction=None):
        links = [links] if isinstance(links, int) else links
        if lateral_friction is not None:
            for l in links:
                self.sim.set_link_friction(self.link0 if l == self.base else self._gl(l), lateral_friction)

    def set_whole_body_frictions(self, lateral_friction=None, spinning_friction=None, rolling_friction=None):
        self.set_frictions(self.all_joint_indices, lateral_friction, spinning_friction, rolling_friction)

    # ---- limits (agent.py:209-250)
    def update_joint_limits(self, indices=None):
        indices = self.all_joint_indices if indices is None else indices
        s = self.sim.scene
        self.lower_limits, self.upper_limits = {}, {}
        ik_lo, ik_hi = [], []
        for j in indices:
            g = self._gl(j)
            lo, hi, jt = float(s['link_lower'][g]), float(s['link_upper'][g]), int(s['link_jtype'][g])
            if lo == 0 and hi == -1:
                lo, hi = -1e10, 1e10
                if jt != 0:
                    ik_lo.append(-2 * np.pi)
                    ik_hi.append(2 * np.pi)
            elif jt != 0:
                ik_lo.append(lo)
                ik_hi.append(hi)
            self.lower_limits[j], self.upper_limits[j] = lo, hi
        self.ik_lower_limits, self.ik_upper_limits = np.array(ik_lo), np.array(ik_hi)

    def enforce_joint_limits(self, indices=None):
        indices = self.all_joint_indices if indices is None else indices
        g = [self._gl(j) for j in indices]
        q, qd, _ = self.sim.get_joint_states(g)
        lo = np.array([self.lower_limits[j] for j in indices])
        hi = np.array([self.upper_limits[j] for j in indices])
        bad = (q < lo) | (q > hi)
        if bad.any():
            self.sim.set_joint_state(g, q=np.clip(q, lo, hi), qd=np.where(bad, 0.0, qd))
