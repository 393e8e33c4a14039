"""`Human` — joint / link index tables of the capsule person (reference envs/agents/human.py:5-58)."""
import numpy as np

from .agent import Agent

right_arm_joints = list(range(0, 10))
left_arm_joints = list(range(10, 20))
head_joints = [20, 21, 22, 23]
right_leg_joints = list(range(28, 35))
left_leg_joints = list(range(35, 42))


class Human(Agent):
    # limb (link) indices, human.py:17-35
    right_pecs, right_shoulder, right_elbow, right_wrist = 2, 5, 7, 9
    left_pecs, left_shoulder, left_elbow, left_wrist = 12, 15, 17, 19
    neck, head, stomach, waist = 20, 23, 24, 27
    right_hip, right_knee, right_ankle, left_hip, left_knee, left_ankle = 30, 31, 34, 37, 38, 41
    # joint indices, human.py:37-55
    j_right_pecs_x, j_right_pecs_y, j_right_pecs_z = 0, 1, 2
    j_right_shoulder_x, j_right_shoulder_y, j_right_shoulder_z = 3, 4, 5
    j_right_elbow, j_right_forearm, j_right_wrist_x, j_right_wrist_y = 6, 7, 8, 9
    j_left_pecs_x, j_left_pecs_y, j_left_pecs_z = 10, 11, 12
    j_left_shoulder_x, j_left_shoulder_y, j_left_shoulder_z = 13, 14, 15
    j_left_elbow, j_left_forearm, j_left_wrist_x, j_left_wrist_y = 16, 17, 18, 19
    j_neck, j_head_x, j_head_y, j_head_z = 20, 21, 22, 23
    j_waist_x, j_waist_y, j_waist_z = 25, 26, 27
    j_right_hip_x, j_right_hip_y, j_right_hip_z, j_right_knee = 28, 29, 30, 31
    j_right_ankle_x, j_right_ankle_y, j_right_ankle_z = 32, 33, 34
    j_left_hip_x, j_left_hip_y, j_left_hip_z, j_left_knee = 35, 36, 37, 38
    j_left_ankle_x, j_left_ankle_y, j_left_ankle_z = 39, 40, 41

    def __init__(self, controllable_joint_indices, controllable=False):
        super().__init__()
        self.controllable_joint_indices = controllable_joint_indices
        self.controllable = controllable
        self.right_arm_joints, self.left_arm_joints, self.head_joints = right_arm_joints, left_arm_joints, head_joints
        self.right_leg_joints, self.left_leg_joints = right_leg_joints, left_leg_joints
        self.impairment, self.limit_scale, self.strength = 'none', 1.0, 1.0
        self.tremors = np.zeros(10)
        self.motor_forces, self.motor_gains = 1.0, 0.05
        self.gender = 'male'          # per-env genders live in the env (`male` mask); N == 1 mirrors the reference attribute
        self.limit_scale_env = None   # [n_envs] per-env scale of the joint limits (impairment 'limits', human.py:85, human_creation.py:199-236) or None
        self.limits_model = None      # the realistic joint-limit classifier (human.py:73), loaded on first use
        self.arm_previous_valid_pose = {True: None, False: None}      # human.py:68: per arm, here [n_envs][4] with NaN = none yet

    def set_limit_scale(self, scale):
        """Per-env joint-limit scale of the `limits` impairment.  The reference bakes it into the person's URDF-like description
        (human_creation.py:199-236: lower and upper limits times the scale); the batched scene template is shared by the envs, so
        here it acts through the host-side clamps: the action clamp of `take_step` (controllable limits become [n_envs][k]) and
        `enforce_joint_limits` after every substep."""
        self.limit_scale_env = np.asarray(scale, dtype=np.float64).reshape(-1)
        lo = np.array([self.lower_limits[i] for i in self.controllable_joint_indices]); hi = np.array([self.upper_limits[i] for i in self.controllable_joint_indices])
        self.controllable_joint_lower_limits = lo[None, :] * self.limit_scale_env[:, None]
        self.controllable_joint_upper_limits = hi[None, :] * self.limit_scale_env[:, None]

    def enforce_joint_limits(self, indices=None):
        if self.limit_scale_env is None:
            return super().enforce_joint_limits(indices)
        indices = self.all_joint_indices if indices is None else indices
        g = [self._gl(j) for j in indices]
        q, qd, _ = self.sim.get_joint_states(g)
        lo = np.array([self.lower_limits[j] for j in indices])[None, :] * self.limit_scale_env[:, None]
        hi = np.array([self.upper_limits[j] for j in indices])[None, :] * self.limit_scale_env[:, None]
        bad = (q < lo) | (q > hi)
        if bad.any():
            self.sim.set_joint_state(g, q=np.clip(q, lo, hi), qd=np.where(bad, 0.0, qd))

    def enforce_realistic_joint_limits(self, env_mask=None):
        """human.py:134-152, for every env at once: the shoulder / elbow angles of the controllable arm are classified by the
        joint-limit MLP; a reachable pose is remembered, an unreachable one is replaced by the env's last reachable pose (joint
        velocities zeroed, as `set_joint_angles` does).  `env_mask`: envs in which this person exists (the other gender's copy
        is switched off)."""
        ci = self.controllable_joint_indices
        if self.j_right_shoulder_x not in ci and self.j_left_shoulder_x not in ci:
            return
        right = self.j_right_shoulder_x in ci
        indices = ([self.j_right_shoulder_x, self.j_right_shoulder_y, self.j_right_shoulder_z, self.j_right_elbow] if right else
                   [self.j_left_shoulder_x, self.j_left_shoulder_y, self.j_left_shoulder_z, self.j_left_elbow])
        if self.limits_model is None:
            from ...limits_model import load_model
            self.limits_model = load_model()
        ang = np.atleast_2d(self.get_joint_angles(indices))
        tz, tx, ty, qe = ang.T
        sgn = -1.0 if right else 1.0
        two_pi = 2 * np.pi
        x = np.stack([(sgn * tz + two_pi) % two_pi, (tx + two_pi) % two_pi, sgn * ty, (-qe + two_pi) % two_pi], axis=1)     # the angle convention of the training data
        ok = self.limits_model.predict_classes(x)[:, 0] == 1
        if env_mask is not None:
            ok = ok | ~np.asarray(env_mask, dtype=bool)
        prev = self.arm_previous_valid_pose[right]
        if prev is None:
            prev = np.full(ang.shape, np.nan)
        prev = np.where(ok[:, None], ang, prev)
        self.arm_previous_valid_pose[right] = prev
        back = ~ok & ~np.isnan(prev[:, 0])
        if back.any():
            lo = np.array([self.lower_limits[j] for j in indices]); hi = np.array([self.upper_limits[j] for j in indices])
            self.sim.set_joint_state([self._gl(j) for j in indices], q=np.clip(np.where(back[:, None], prev, ang), lo, hi), qd=np.zeros_like(ang),
                                     mask=back.astype(np.int32))
            self.sim.forward_kinematics()
