"""`FeedingEnv` (reference envs/feeding.py) on the batched backend.

`step` runs the fused kernels (`ag_feeding_step_host`): action -> PD targets -> 5 substeps -> obs /
reward / done.  `step_reference_api` performs the same step the way the reference does it —
`take_step` + `_get_obs` + `get_food_rewards` + `human_preferences` through the per-call `Agent`
API — and exists so that tests can show the two paths agree."""
import numpy as np

from .. import capi
from ..feeding_batch import HEAD_LINK, IMPAIRMENTS, FeedingBatch
from ..kinematics import q_rot
from ..sim import BatchSim
from .agents.agent import Agent
from .agents.furniture import Furniture
from .env import AssistiveEnv


class FeedingEnv(AssistiveEnv):
    def __init__(self, robot, human, n_envs=1, device=0, seed=1001, config=None):
        super().__init__(robot=robot, human=human, task='feeding', n_envs=n_envs, device=device, seed=seed,
                         obs_robot_len=(18 + len(robot.controllable_joint_indices) - (len(robot.wheel_joint_indices) if robot.mobile else 0)),
                         obs_human_len=(19 + len(human.controllable_joint_indices)))
        self._fb = FeedingBatch()
        self.human_impairment = 'random'      # build_assistive_env(human_impairment='random'), env.py:114
        self._cfg = config or capi.default_config()
        self._sim_lib = None
        self.total_food_count = 8

    # ------------------------------------------------------------------ reset (feeding.py:114-182)
    def reset(self):
        super().reset()
        fb = self._fb
        if self.id is None:
            self.id = BatchSim(fb.scene, self._cfg, self.n_envs, device=self.device, _lib=self._sim_lib)
            sim = self.id
            self.plane.init(fb.plane, sim, self.np_random, indices=-1)
            self.robot.init(fb.robot, sim, self.np_random)
            self.tool.init(fb.tool, sim, self.np_random, indices=-1)
            self.furniture.init(fb.wheelchair, sim, self.np_random, indices=-1)
            self.table, self.bowl = Furniture(), Furniture()
            self.table.init(fb.table, sim, self.np_random, indices=-1)
            self.bowl.init(fb.bowl, sim, self.np_random, indices=-1)
            self.humans = {}
            for g, hb in fb.humans.items():
                h = type(self.human)(self.human.controllable_joint_indices, controllable=self.human.controllable)
                h.init(hb, sim, self.np_random, self.human.controllable_joint_indices)
                self.humans[g] = h
            self.foods_agents = []
            for f in fb.foods:
                a = Agent()
                a.init(f, sim, self.np_random, indices=-1)
                self.foods_agents.append(a)
            self._feeding_ready = False
        rng = np.random.default_rng(self.np_random.randint(0, 2 ** 31 - 1))
        self.robot.motor_gains = self.human.motor_gains = 0.025          # feeding.py:122
        self.agents = [self.robot]
        coop = bool(self.human.controllable)
        # (co-optimisation: the reference may also draw `tremor` for a controllable person, human.py:80-81; not combined here)
        s = fb.reset(self.id, rng, settle_steps=25, impairment='no_tremor' if coop else self.human_impairment, simulate_head=coop)
        self.male = s['male'].astype(bool)
        self.human.gender = 'male' if self.male[0] else 'female'
        # impairments (human.py:79-92).  The reference appends a tremor human to `agents`
        # (env.py:130-131); here each gender's Human carries a per-env tremor mask.
        tremor = s['impairment'] == 3
        self.impairment = s['impairment']
        self.human.impairment = IMPAIRMENTS[int(s['impairment'][0])]
        self.human.limit_scale, self.human.strength = float(s['limit_scale'][0]), float(s['strength'][0])
        rest = fb.tremor_rest_of(s)
        for g, h in self.humans.items():
            h.tremor_mask = tremor & (self.male if g == 'male' else ~self.male)
            h.impairment = 'tremor' if h.tremor_mask.any() else 'none'
            h.tremors = np.where(h.tremor_mask[:, None], s['tremors'], 0.0)
            h.target_joint_angles = rest
            h.motor_gains = self.human.motor_gains
            if h.tremor_mask.any():
                self.agents.append(h)
        if coop:                                  # both gender instances act; the switched-off one moves nothing (env.py:130)
            for h in self.humans.values():
                h.motor_gains, h.motor_forces = self.human.motor_gains, self.human.motor_forces
                self.agents.append(h)
        self.mouth_pos = np.where(self.male[:, None], fb.mouth['male'], fb.mouth['female'])
        if not self._feeding_ready:
            fb.start_fused(self.id, s, seed=self._seed)
            self._feeding_ready = True
        else:
            fb.start_fused(self.id, s, seed=self._seed)
        self.foods = np.ones((self.n_envs, 8), dtype=bool)
        self.foods_active = np.ones((self.n_envs, 8), dtype=bool)
        self.task_success = np.zeros(self.n_envs, dtype=int)
        self.update_targets()
        return self._squeeze(self._get_obs())

    def _squeeze(self, a):
        if isinstance(a, dict):
            return {k: self._squeeze(v) for k, v in a.items()}
        return a[0] if self.n_envs == 1 else a

    # ------------------------------------------------------------------ fused step (feeding.py:12-43)
    def step(self, action):
        if self.human.controllable:               # feeding.py:13-14,40-43: dict in, dicts out (per-call API path)
            a = np.concatenate([np.asarray(action['robot'], dtype=np.float64).reshape(self.n_envs, -1), np.asarray(action['human'], dtype=np.float64).reshape(self.n_envs, -1)], axis=1)
            obs, reward, done, info = self.step_reference_api(a)
            d = bool(np.all(done)) if self.n_envs > 1 else bool(done)
            return obs, {'robot': reward, 'human': reward}, {'robot': done, 'human': done, '__all__': d}, {'robot': info, 'human': info}
        a = np.asarray(action, dtype=np.float32).reshape(self.n_envs, -1)
        obs, rew, done, info = self.id.feeding_step_host(a)
        self.iteration += 1
        self.total_force_on_human = info[:, 0]
        infos = [{'total_force_on_human': float(info[e, 0]), 'task_success': int(info[e, 1]), 'action_robot_len': self.action_robot_len,
                  'action_human_len': self.action_human_len, 'obs_robot_len': self.obs_robot_len, 'obs_human_len': self.obs_human_len}
                 for e in range(self.n_envs)]
        if self.n_envs == 1:
            return obs[0], float(rew[0]), bool(done[0] > 0.5), infos[0]
        return obs, rew, done > 0.5, infos

    # ------------------------------------------------------------------ the same step through the reference-shaped API
    def _head_pose(self):
        pm, qm = self.humans['male'].get_pos_orient(HEAD_LINK)
        pf, qf = self.humans['female'].get_pos_orient(HEAD_LINK)
        pm, qm, pf, qf = (np.atleast_2d(x) for x in (pm, qm, pf, qf))
        return np.where(self.male[:, None], pm, pf), np.where(self.male[:, None], qm, qf)

    def update_targets(self):                                            # feeding.py:192-196
        hp, hq = self._head_pose()
        self.target_pos = hp + q_rot(hq, self.mouth_pos)

    def get_total_force(self):                                           # feeding.py:45-48
        r = sum(self.id.contact_force_sum(self.robot.body, h.body) for h in self.humans.values())
        s = sum(self.id.contact_force_sum(self.tool.body, h.body) for h in self.humans.values())
        return r.astype(np.float64), s.astype(np.float64)

    def _get_obs(self, agent=None):                                      # feeding.py:85-112
        sp, sq = (np.atleast_2d(x) for x in self.tool.get_base_pos_orient())
        sp_r, sq_r = (np.atleast_2d(x) for x in self.robot.convert_to_realworld(sp, sq))
        q = np.atleast_2d(self.robot.get_joint_angles(self.robot.controllable_joint_indices))
        q = (q + np.pi) % (2 * np.pi) - np.pi
        hp, hq = self._head_pose()
        hp_r, hq_r = (np.atleast_2d(x) for x in self.robot.convert_to_realworld(hp, hq))
        tg_r = np.atleast_2d(self.robot.convert_to_realworld(self.target_pos)[0])
        self.robot_force_on_human, self.spoon_force_on_human = self.get_total_force()
        self.total_force_on_human = self.robot_force_on_human + self.spoon_force_on_human
        robot_obs = np.concatenate([sp_r, sq_r, sp_r - tg_r, q, hp_r, hq_r, self.spoon_force_on_human[:, None]], axis=1)
        if agent == 'robot' or not self.human.controllable:
            return robot_obs
        # feeding.py:101-111: the same quantities in the person's base frame + the person's joint angles
        def human_frame(pos, orient=None):
            outs = []
            for g in ('male', 'female'):
                r = self.humans[g].convert_to_realworld(pos, orient if orient is not None else np.array([0, 0, 0, 1.0]))
                outs.append([np.atleast_2d(x) for x in r])
            return [np.where(self.male[:, None], m, f) for m, f in zip(*outs)]
        qh = np.where(self.male[:, None], np.atleast_2d(self.humans['male'].get_joint_angles(self.human.controllable_joint_indices)),
                      np.atleast_2d(self.humans['female'].get_joint_angles(self.human.controllable_joint_indices)))
        sp_h, sq_h = human_frame(sp, sq)
        hp_h, hq_h = human_frame(hp, hq)
        tg_h = human_frame(self.target_pos)[0]
        human_obs = np.concatenate([sp_h, sq_h, sp_h - tg_h, qh, hp_h, hq_h, self.robot_force_on_human[:, None], self.spoon_force_on_human[:, None]], axis=1)
        if agent == 'human':
            return human_obs
        return {'robot': robot_obs, 'human': human_obs}

    def get_food_rewards(self):                                          # feeding.py:50-83
        n = self.n_envs
        food_reward, hit_reward, vel_sum = np.zeros(n), np.zeros(n), np.zeros(n)
        active_entry = self.foods_active.copy()
        for i, f in enumerate(self.foods_agents):
            fp = np.atleast_2d(f.get_base_pos_orient()[0])
            dist = np.linalg.norm(self.target_pos - fp, axis=1)
            near = self.id.closest_points(f.body, self.tool.body, 0.1, max_pts=1)[1] > 0
            eaten = self.foods[:, i] & (dist < 0.03)
            spilled = self.foods[:, i] & ~eaten & ~near
            food_reward += 20.0 * eaten - 5.0 * spilled
            self.task_success += eaten
            vel_sum += eaten * np.linalg.norm(np.atleast_2d(f.get_velocity(f.base)), axis=1)
            self.foods[:, i] &= ~(eaten | spilled)
            self.foods_active[:, i] &= ~eaten
            if eaten.any():   # teleport eaten food far away (feeding.py:69)
                far = self.np_random.uniform(1000, 2000, size=(n, 3))
                self.id.set_base_pose(f.body, np.where(eaten[:, None], far, fp), None, mask=eaten.astype(np.int32))
        for i, f in enumerate(self.foods_agents):
            touching = sum(self.id.get_contacts(f.body, h.body, max_pts=1)[1] for h in self.humans.values()) > 0
            hit = active_entry[:, i] & touching
            hit_reward -= hit
            self.foods_active[:, i] &= ~hit
        return food_reward, vel_sum, hit_reward

    def step_reference_api(self, action):
        a = np.asarray(action, dtype=np.float64).reshape(self.n_envs, -1)
        self.take_step(a)
        obs = self._get_obs()
        reward_food, vel_sum, food_hit = self.get_food_rewards()
        ee_vel = np.linalg.norm(np.atleast_2d(self.robot.get_velocity(self.robot.right_end_effector)), axis=1)
        pref = (self.C_v * (-ee_vel) + self.C_f * (-self.total_force_on_human) +
                self.C_hf * np.where(self.spoon_force_on_human < 10, 0.0, -self.spoon_force_on_human) + self.C_fd * food_hit + self.C_fdv * (-vel_sum))
        spoon_pos = np.atleast_2d(self.tool.get_base_pos_orient()[0])
        reward = (self.config('distance_weight') * (-np.linalg.norm(self.target_pos - spoon_pos, axis=1)) +
                  self.config('action_weight') * (-np.linalg.norm(a, axis=1)) + self.config('food_reward_weight') * reward_food + pref)
        done = np.full(self.n_envs, self.iteration >= 200)
        info = {'total_force_on_human': self.total_force_on_human, 'task_success': (self.task_success >= self.total_food_count * self.config('task_success_threshold')).astype(int),
                'action_robot_len': self.action_robot_len, 'action_human_len': self.action_human_len, 'obs_robot_len': self.obs_robot_len, 'obs_human_len': self.obs_human_len}
        return self._squeeze(obs), self._squeeze(reward), self._squeeze(done), info
