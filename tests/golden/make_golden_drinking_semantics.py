"""Golden rollout of the reference's OWN `DrinkingEnv.step` (envs/drinking.py:10-117 + env.py:174-274, unmodified) executed on the CPU
oracle through the pybullet facade of make_golden_feeding_semantics.py: the cup with 64 water particles, `numSubSteps = 4`,
`numSolverIterations = 10`.  The robot tips the cup (wrist joint) so that particles leave it and are counted as spilled, and one
particle is tossed up from the mouth target so that it is counted as swallowed.  Output: tests/golden/drinking_semantics.npz,
replayed by tests/test_reference_drinking_semantics.py with `DrinkingJacoEnv` of this repo on the oracle.

usage: python tests/golden/make_golden_drinking_semantics.py [/root/reference]"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

N_STEPS, SEED, SWALLOW_STEP, SPILL_STEP = 24, 0, 6, 14
# symplectic Euler over 5 x 4 substeps of 0.005 s: back at the start height after one env step if 0.1 v0 = g dt^2 (1 + ... + 20)
SWALLOW_V0 = 9.81 * 0.005 ** 2 * 210 / 0.1


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    from assistive_gym_b200.drinking_batch import DrinkingBatch
    from oracle.oracle_py import OracleSim
    db = DrinkingBatch()
    sim = OracleSim(db.scene, DrinkingBatch.config(), 1)
    smp = db.reset(sim, np.random.default_rng(SEED), settle_steps=50, impairment='none')
    start_state = sim.state_get()
    male = bool(smp['male'][0])
    hb = db.humans['male' if male else 'female']
    from make_golden_env_logic import install_stubs
    from make_golden_feeding_semantics import Facade
    install_stubs(ref)
    import assistive_gym  # noqa: F401  (the reference package)
    from assistive_gym.envs.agents.agent import Agent
    from assistive_gym.envs.drinking_envs import DrinkingJacoEnv
    from assistive_gym.envs.util import Util
    env = DrinkingJacoEnv()
    p = sys.modules['pybullet']
    Facade(sim, db.scene, f32_targets=True).install(p)

    def euler(q, physicsClientId=None):                                 # PyBullet's getEulerFromQuaternion: roll, pitch, yaw (fixed axes x, y, z)
        x, y, z, w = [float(v) for v in q]
        sinp = 2 * (w * y - z * x)
        return (np.arctan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y)), np.arcsin(np.clip(sinp, -1, 1)), np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z)))
    p.getEulerFromQuaternion = euler

    def quat(e, physicsClientId=None):
        r, pt, y = [float(v) for v in e]
        cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(pt / 2), np.sin(pt / 2), np.cos(y / 2), np.sin(y / 2)
        return [sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy]
    p.getQuaternionFromEuler = quat
    env.robot.body, env.tool.body, env.human.body = db.robot, db.tool, hb
    env.human.gender = 'male' if male else 'female'
    for a in (env.robot, env.tool, env.human):
        a.id = 0
    env.robot.controllable_joint_lower_limits = np.array(db.arm_lower, dtype=np.float64)
    env.robot.controllable_joint_upper_limits = np.array(db.arm_upper, dtype=np.float64)
    env.robot.motor_gains = env.human.motor_gains = 0.005               # drinking.py:126
    env.agents = [env.robot]
    env.waters = []
    for w in db.waters:
        a = Agent(); a.body, a.id = w, 0
        env.waters.append(a)
    env.waters_active = list(env.waters)
    env.total_water_count = len(env.waters)
    env.cup_top_center_offset, env.cup_bottom_center_offset = np.array([0, 0, -0.055]), np.array([0, 0, 0.07])      # drinking.py:137-138
    env.mouth_pos = [0, -0.11, 0.03] if male else [0, -0.1, 0.03]
    env.target = types.SimpleNamespace(set_base_pos_orient=lambda *a, **k: None)
    env.iteration, env.task_success, env.last_sim_time, env.gui = 0, 0, None, False
    env.action_space = types.SimpleNamespace(low=-np.ones(7), high=np.ones(7))
    env.np_random = np.random.RandomState(0)
    if getattr(env, 'util', None) is None:
        env.util = Util(0, env.np_random)
    env.update_targets()
    actions = np.zeros((N_STEPS, 7))
    actions[:, 5] = 1.0                                                 # tip the cup with the wrist ...
    actions[:, 6] = 1.0
    actions[:, :5] = np.random.default_rng(SEED + 1).uniform(-1, 1, size=(N_STEPS, 5)) * 0.2
    obs, rew, done, total, success, n_w, n_a = [], [], [], [], [], [], []
    swallow, swallow_pos, spill, spill_pos = -1, np.zeros(3), -1, np.zeros(3)
    for t in range(N_STEPS):
        if t == SWALLOW_STEP:
            swallow = db.waters.index(env.waters[0].body)
            head = sim.get_link_states([db.gl(hb, 23)])['pos'][0, 0]
            swallow_pos = env.target_pos + 0.015 * (env.target_pos - head) / np.linalg.norm(env.target_pos - head)      # 1.5 cm in front of the mouth: clear of the face
            sim.set_base_pose(db.waters[swallow], swallow_pos[None], np.array([[0, 0, 0, 1.0]]))
            sim.set_base_velocity(db.waters[swallow], np.array([[0, 0, SWALLOW_V0]]), np.zeros((1, 3)))
        if t == SPILL_STEP:                                                # forced event: a particle appears 30 cm beside the cup (out of it, farther than 0.1 m)
            spill = db.waters.index(env.waters[0].body)
            spill_pos = sim.get_link_states([int(db.scene['body_link0'][db.tool])])['com_pos'][0, 0] + np.array([0.3, 0.0, 0.0])
            sim.set_base_pose(db.waters[spill], spill_pos[None], np.array([[0, 0, 0, 1.0]]))
            sim.set_base_velocity(db.waters[spill], np.zeros((1, 3)), np.zeros((1, 3)))
        o, r, d, info = env.step(actions[t].copy())
        obs.append(np.asarray(o, dtype=np.float64)); rew.append(float(r)); done.append(bool(d)); total.append(float(info['total_force_on_human']))
        success.append(int(env.task_success)); n_w.append(len(env.waters)); n_a.append(len(env.waters_active))
    out = {('sample_' + k): np.asarray(v) for k, v in smp.items()}
    out.update(start_state=start_state, actions=actions, obs=np.array(obs), reward=np.array(rew), done=np.array(done), total_force=np.array(total), task_success=np.array(success),
               n_waters=np.array(n_w), n_waters_active=np.array(n_a), swallow_step=np.array(SWALLOW_STEP), swallow_water=np.array(swallow), swallow_v0=np.array(SWALLOW_V0), swallow_pos=np.array(swallow_pos), spill_step=np.array(SPILL_STEP), spill_water=np.array(spill), spill_pos=np.array(spill_pos), seed=np.array(SEED))
    np.savez_compressed(os.path.join(HERE, 'drinking_semantics.npz'), **out)
    print('waters left', n_w, 'active', n_a[-1], 'swallowed', success[-1], 'reward', np.round(rew, 2))


if __name__ == '__main__':
    main()
