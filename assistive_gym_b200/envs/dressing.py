"""`DressingEnv` (reference envs/dressing.py) on the batched backend.

`step` runs the fused path (`ag_dressing_step_host`): action -> PD targets -> 5 x (8 rigid substeps + one cloth launch, the
cloth's anchor follows the end effector) -> sleeve-on-arm reward, cloth forces, obs[24] / reward / done.  `_get_obs` (used
by `reset`) reads the same quantities through the per-call Agent API."""
import numpy as np

from .. import capi
from ..dressing_batch import L_ELBOW, L_SHOULDER, L_WRIST, DressingBatch
from ..sim import BatchSim
from .env import AssistiveEnv


class DressingEnv(AssistiveEnv):
    def __init__(self, robot, human, n_envs=1, device=0, seed=1001, config=None, toc_attempts=50):
        super().__init__(robot=robot, human=human, task='dressing', n_envs=n_envs, device=device, seed=seed,
                         obs_robot_len=(17 + len(robot.controllable_joint_indices) - (len(robot.wheel_joint_indices) if robot.mobile else 0)),
                         obs_human_len=(18 + len(human.controllable_joint_indices)))
        self._db = DressingBatch()
        self._cfg = config or DressingBatch.config()                       # numSubSteps = 8 (dressing.py:184)
        self._toc_attempts = toc_attempts
        self._sim_lib = None

    def step(self, action):                                                # dressing.py:12-77
        a = np.asarray(action, dtype=np.float32).reshape(self.n_envs, -1)
        obs, rew, done, info = self.id.dressing_step_host(a)
        self.iteration += 1
        self.total_force_on_human, self.cloth_force_sum = info[:, 0], obs[:, 23]
        self.task_success = np.maximum(self.task_success, info[:, 2])
        self.forearm_in_sleeve, self.upperarm_in_sleeve = (info[:, 3].astype(int) & 1) > 0, (info[:, 3].astype(int) & 2) > 0
        out = {'total_force_on_human': info[:, 0], 'task_success': info[:, 1].astype(int), 'action_robot_len': self.action_robot_len,
               'action_human_len': self.action_human_len, 'obs_robot_len': self.obs_robot_len, 'obs_human_len': self.obs_human_len}
        if self.n_envs == 1:
            return obs[0], float(rew[0]), bool(done[0] > 0.5), {k_: (v[0] if isinstance(v, np.ndarray) else v) for k_, v in out.items()}
        return obs, rew, done > 0.5, out

    def _arm_points(self):
        out = []
        for link in (L_SHOULDER, L_ELBOW, L_WRIST):
            pm = np.atleast_2d(self.humans['male'].get_pos_orient(link)[0])
            pf = np.atleast_2d(self.humans['female'].get_pos_orient(link)[0])
            out.append(np.where(self.male[:, None], pm, pf))
        return out

    def _get_obs(self, agent=None):                                        # dressing.py:79-106
        ep, eq = (np.atleast_2d(x) for x in self.robot.get_pos_orient(self.robot.left_end_effector))
        ep_r, eq_r = (np.atleast_2d(x) for x in self.robot.convert_to_realworld(ep, eq))
        q = np.atleast_2d(self.robot.get_joint_angles(self.robot.controllable_joint_indices))
        q = (q + np.pi) % (2 * np.pi) - np.pi
        arm = [np.atleast_2d(self.robot.convert_to_realworld(p_)[0]) for p_ in self._arm_points()]
        cnt, _node, pos, force, _link = self.id.cloth_get_contacts(1024)
        f = np.linalg.norm(force * 10.0, axis=2)
        keep = (np.arange(f.shape[1])[None, :] < cnt[:, None]) & (pos[:, :, 2] < ep[:, 2:3] - 0.05) & (f < 20)
        self.cloth_force_sum = np.where(keep, f, 0.0).sum(axis=1)
        self.robot_force_on_human = sum(self.id.contact_force_sum(self.robot.body, h.body) for h in self.humans.values()).astype(np.float64)
        self.total_force_on_human = self.robot_force_on_human + self.cloth_force_sum
        return np.concatenate([ep_r, eq_r, q] + arm + [self.cloth_force_sum[:, None]], axis=1)

    def reset(self):                                                       # dressing.py:108-198
        super().reset()
        db = self._db
        if self.id is None:
            self.id = BatchSim(db.scene, self._cfg, self.n_envs, device=self.device, _lib=self._sim_lib)
            sim = self.id
            self.plane.init(db.plane, sim, self.np_random, indices=-1)
            self.robot.init(db.robot, sim, self.np_random)
            self.furniture.init(db.wheelchair, sim, self.np_random, indices=-1)
            self.humans = {}
            for g, hb in db.humans.items():
                h = type(self.human)(self.human.controllable_joint_indices, controllable=False)
                h.init(hb, sim, self.np_random, self.human.controllable_joint_indices)
                self.humans[g] = h
        rng = np.random.default_rng(self.np_random.randint(0, 2 ** 31 - 1))
        self.agents = [self.robot]
        self.robot.motor_gains = self.human.motor_gains = 0.01             # dressing.py:117
        s = db.reset(self.id, rng, attempts=self._toc_attempts)
        self.male = s['male'].astype(bool)
        self.human.gender = 'male' if self.male[0] else 'female'
        self.start_ee_pos = db.start_ee_pos
        db.start_fused(self.id, s)
        self.task_success = np.zeros(self.n_envs)
        obs = self._get_obs()
        return obs[0] if self.n_envs == 1 else obs

    def update_targets(self):                                              # dressing.py:200-210
        self.id.cloth_anchor_follow(self._db.ee_link)
