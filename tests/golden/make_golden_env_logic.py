"""Golden vectors of the reference's env-level logic, recorded by importing the reference package in this container with every
third-party module it imports (pybullet, gym, keras, ray, screeninfo, smplx ...) replaced by an inert stub and calling its own
methods:
  * `AssistiveEnv.take_step(..., step_sim=False)` (envs/env.py:174-222): action clip / scale, the 5-fold accumulation with the
    joint-limit clamp -> the motor targets handed to `agent.control` (the agent's measured angles and limits are injected);
  * `AssistiveEnv.human_preferences` (envs/env.py:237-274) for the four built tasks on random inputs;
  * the per-task reward weights of config.ini as `env.config(...)` returns them, and the robot classes' constants
    (agents/jaco.py, sawyer.py, pr2.py: joint index tables, tool offsets, gripper positions, base offsets).
Written to tests/golden/env_logic.json; tests/test_reference_env_logic.py compares the repo's mirrors against it.

usage: python tests/golden/make_golden_env_logic.py [/root/reference]"""
import json
import os
import sys
import types

import numpy as np


class Z(int):
    def __call__(self, *a, **k):
        return Z(0)


class Any(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return Z(0)


def install_stubs(ref):
    for name in ['gym', 'gym.spaces', 'gym.utils', 'gym.envs', 'gym.envs.registration', 'screeninfo', 'numpngw', 'keras', 'keras.models', 'ray', 'ray.rllib',
                 'ray.rllib.env', 'ray.rllib.env.multi_agent_env', 'ray.tune', 'ray.tune.registry', 'tensorflow', 'cv2', 'pybullet_data', 'matplotlib',
                 'matplotlib.pyplot', 'smplx', 'trimesh', 'pybullet']:
        sys.modules[name] = Any(name)
    g = sys.modules['gym']
    g.spaces, g.utils, g.Env = sys.modules['gym.spaces'], sys.modules['gym.utils'], object
    sys.modules['gym.utils'].seeding = types.SimpleNamespace(np_random=lambda seed=None: (np.random.RandomState(seed), seed))
    sys.modules['gym.envs.registration'].register = lambda **k: None
    sys.modules['keras.models'].load_model = lambda *a, **k: None
    sys.modules['screeninfo'].get_monitors = lambda: []
    sys.modules['ray.rllib.env.multi_agent_env'].MultiAgentEnv = object
    sys.modules['ray.tune.registry'].register_env = lambda *a, **k: None
    sys.modules['gym.spaces'].Box = lambda low=None, high=None, dtype=None, **k: types.SimpleNamespace(low=np.asarray(low), high=np.asarray(high))
    sys.path.insert(0, ref)


def jsonable(v):
    if isinstance(v, dict):
        return {str(k): jsonable(x) for k, x in v.items()}
    if isinstance(v, (list, tuple, np.ndarray)):
        return [jsonable(x) for x in v]
    if isinstance(v, (np.floating, float)):
        return float(v)
    if isinstance(v, (np.integer, int, bool)):
        return int(v)
    return v if v is None or isinstance(v, str) else str(v)


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    install_stubs(ref)
    import assistive_gym  # noqa: F401  (the reference package)
    from assistive_gym.envs.bed_bathing_envs import BedBathingSawyerEnv
    from assistive_gym.envs.dressing_envs import DressingPR2Env
    from assistive_gym.envs.feeding_envs import FeedingJacoEnv
    from assistive_gym.envs.scratch_itch_envs import ScratchItchJacoEnv
    rng = np.random.default_rng(0)
    out = {'tasks': {}}
    for cls in (FeedingJacoEnv, BedBathingSawyerEnv, DressingPR2Env, ScratchItchJacoEnv):
        env = cls()
        t = {'task': env.task, 'obs_robot_len': env.obs_robot_len, 'obs_human_len': env.obs_human_len, 'action_robot_len': env.action_robot_len,
             'weights': {k: float(env.config(k)) for k in env.configp[env.task]},
             'preference_weights': dict(C_v=env.C_v, C_f=env.C_f, C_hf=env.C_hf, C_fd=env.C_fd, C_fdv=env.C_fdv, C_d=env.C_d, C_p=env.C_p)}
        robot = env.robot
        t['robot'] = {k: jsonable(v) for k, v in vars(robot).items() if k in (
            'controllable_joint_indices', 'right_arm_joint_indices', 'left_arm_joint_indices', 'right_end_effector', 'left_end_effector', 'right_gripper_indices',
            'left_gripper_indices', 'gripper_pos', 'right_tool_joint', 'left_tool_joint', 'tool_pos_offset', 'tool_orient_offset', 'right_gripper_collision_indices',
            'left_gripper_collision_indices', 'toc_base_pos_offset', 'toc_ee_orient_rpy', 'wheelchair_mounted', 'motor_forces', 'motor_gains', 'action_multiplier')}
        # ---- human_preferences on random inputs
        prefs = []
        for _ in range(40):
            kw = dict(end_effector_velocity=float(rng.uniform(0, 1)), total_force_on_human=float(rng.uniform(0, 30)), tool_force_at_target=float(rng.uniform(0, 20)),
                      food_hit_human_reward=int(-rng.integers(0, 3)), food_mouth_velocities=[float(v) for v in rng.uniform(0, 1, size=int(rng.integers(0, 3)))],
                      dressing_forces=[[float(x) for x in v] for v in rng.normal(size=(int(rng.integers(1, 4)), 3))])
            prefs.append(dict(kw, out=float(env.human_preferences(**kw))))
        t['human_preferences'] = prefs
        # ---- take_step: the targets it hands to agent.control
        n = len(robot.controllable_joint_indices)
        lo, hi = rng.uniform(-3, -0.5, size=n), rng.uniform(0.5, 3, size=n)
        robot.controllable_joint_lower_limits, robot.controllable_joint_upper_limits = lo, hi
        env.agents = [robot]
        env.iteration, env.last_sim_time = 0, None
        env.action_space = types.SimpleNamespace(low=-np.ones(n), high=np.ones(n))
        steps = []
        for _ in range(60):
            q = rng.uniform(lo - 0.05, hi + 0.05)                       # also measured angles slightly outside the limits
            if rng.random() < 0.5:
                q = np.where(rng.random(n) < 0.4, np.where(rng.random(n) < 0.5, lo + 0.03, hi - 0.03), q)     # near a limit: the clamp acts
            a = rng.uniform(-1.5, 1.5, size=n)
            got = {}
            robot.get_joint_angles = lambda idx, q=q: np.array(q)
            robot.control = lambda idx, tgt, gains, forces, got=got: got.update(indices=list(idx), targets=np.array(tgt).tolist(), gains=gains, forces=forces)
            env.take_step(np.array(a), step_sim=False)
            steps.append(dict(q=q.tolist(), action=a.tolist(), **jsonable(got)))
        t['take_step'] = dict(lower=lo.tolist(), upper=hi.tolist(), steps=steps)
        out['tasks'][cls.__name__] = jsonable(t)
    # ---- the base-pose ranking of Robot.position_robot_toc (agents/robot.py:173-186, :223-235): the joint-limit weighting method and the two
    # source lines that turn J and the weights into the JLWKI score, executed as they stand in the reference file
    import inspect
    from assistive_gym.envs.agents.jaco import Jaco
    rob = Jaco('right')
    src = inspect.getsource(type(rob).position_robot_toc).split('\n')
    formula = [ln.strip() for ln in src if ln.strip().startswith('det = ') or ln.strip().startswith('jlwki = ')]
    assert len(formula) == 2
    cases = []
    for _ in range(40):
        lo, hi = rng.uniform(-3, -0.5, size=7), rng.uniform(0.5, 3, size=7)
        q = rng.uniform(lo - 0.1, hi + 0.1)
        if rng.random() < 0.5:
            q = np.where(rng.random(7) < 0.4, np.where(rng.random(7) < 0.5, lo + 0.02, hi - 0.02), q)
        J = rng.normal(size=(6, 7))
        W = rob.joint_limited_weighting(q, lo, hi)
        ns = dict(np=np, J=J, joint_limit_weight=W, a=6)
        for ln in formula:
            exec(ln, ns)
        cases.append(dict(q=q.tolist(), lower=lo.tolist(), upper=hi.tolist(), J=J.tolist(), weights=np.diag(W).tolist(), jlwki=float(ns['jlwki'])))
    out['jlwki'] = cases
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'env_logic.json')
    json.dump(out, open(path, 'w'), separators=(',', ':'))
    print('wrote', path, {k: (v['task'], v['obs_robot_len'], len(v['take_step']['steps'])) for k, v in out['tasks'].items()})


if __name__ == '__main__':
    main()
