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

    # ---- the rest of the reference's surface (agent.py:94-98,132-207,252-283)
    def get_joint_max_force(self, indices=None):
        """agent.py:94-98 (getJointInfo[10]: the URDF effort limit)"""
        indices = self.all_joint_indices if indices is None else indices
        mf = self.sim.scene.d.get('link_max_force')
        return [float(mf[self._gl(j)]) if mf is not None else 0.0 for j in indices]

    def get_heights(self, set_on_ground=False):
        """agent.py:132-143: height and base height from the AABBs of the base and every link"""
        mn, mx = self.sim.get_link_aabb([self.link0] + [self._gl(j) for j in self.all_joint_indices])
        ok = mn[..., 2] <= mx[..., 2]                                 # links without colliders have no box
        zmin = np.where(ok, mn[..., 2], np.inf).min(axis=1)
        zmax = np.where(ok, mx[..., 2], -np.inf).max(axis=1)
        height = zmax - zmin
        base_height = np.atleast_2d(self.get_pos_orient(self.base)[0])[:, 2] - zmin
        if set_on_ground:
            self.set_on_ground(base_height)
        return self._sq(height), self._sq(base_height)

    def set_on_ground(self, base_height=None):
        """agent.py:158-162"""
        if base_height is None:
            _, base_height = self.get_heights()
        pos, orient = (np.atleast_2d(a) for a in self.get_base_pos_orient())
        pos = pos.copy()
        pos[:, 2] = np.atleast_1d(base_height) + 0.01
        self.set_base_pos_orient(pos, orient)

    def set_friction(self, links, friction):
        self.set_frictions(links, lateral_friction=friction, spinning_friction=0, rolling_friction=0)

    def set_gravity(self, ax=0.0, ay=0.0, az=-9.81):
        """agent.py:196-197 (`p.setGravity(..., body=self.body)`, a fork feature): the same for every env"""
        self.sim.set_body_gravity(self.body, [ax, ay, az])

    def ik(self, target_joint, target_pos, target_orient, ik_indices, max_iterations=1000, half_range=False, use_current_as_rest=False, randomize_limits=False):
        """agent.py:252-274 (`p.calculateInverseKinematics` from a random rest pose) on the device: damped least squares for the
        joints `ik_indices` -- here the pybullet joint indices to solve for -- from one random start inside the limits;
        `target_orient` None = position only.  Returns the joint angles [n_envs, len(ik_indices)] (squeezed for one env)."""
        if target_orient is None:
            tq = np.full(4, np.nan)
        else:
            tq = np.asarray(target_orient, dtype=np.float64)
            if tq.shape[-1] == 3:                                        # euler angles (agent.py:253-254)
                tq = self.get_quaternion(tq)
        q, _err = self.sim.ik_solve([self._gl(j) for j in ik_indices], self._gl(target_joint), np.broadcast_to(np.asarray(target_pos, dtype=np.float64), (self.sim.n, 3)),
                                    tq, max_restarts=1, iters=min(int(max_iterations), 200), threshold=1e-4, seed=int(self.np_random.randint(1, 2 ** 31 - 1)))
        return self._sq(q.astype(np.float64))

    def print_joint_info(self, show_fixed=True):
        """agent.py:276-283"""
        s = self.sim.scene
        for j in self.all_joint_indices:
            g = self._gl(j)
            if show_fixed or int(s['link_jtype'][g]) != 0:
                print(j, {0: 'fixed', 1: 'revolute', 2: 'prismatic'}.get(int(s['link_jtype'][g])), float(s['link_lower'][g]), float(s['link_upper'][g]))

    def _template_level(self, what):
        raise NotImplementedError('%s changes the scene template, which is immutable once the batched simulation exists: do it on the '
                                  'SceneBuilder before `finalize()` (assistive_gym_b200/scene.py)' % what)

    def set_mass(self, link, mass):                                   # agent.py:185-186 changeDynamics(mass=)
        self._template_level('set_mass')

    def set_joint_stiffness(self, joint, stiffness):                  # agent.py:192-194
        self._template_level('set_joint_stiffness')

    def set_all_joints_stiffness(self, stiffness):                    # agent.py:188-190
        self._template_level('set_all_joints_stiffness')

    def create_constraint(self, parent_link, child, child_link, **kw):   # agent.py:202-207 p.createConstraint
        self._template_level('create_constraint')

    def enable_force_torque_sensor(self, joint):                      # agent.py:199-200
        self._template_level('enable_force_torque_sensor')

    def get_force_torque_sensor(self, joint):                         # agent.py:145-146
        self._template_level('get_force_torque_sensor')

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
