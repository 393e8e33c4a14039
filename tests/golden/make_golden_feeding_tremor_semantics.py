"""Golden rollout of the reference's OWN Feeding step with a person who has the `tremor` impairment (envs/env.py:130-131: such a person
joins `agents`; :212-215: its head joints are driven to `target_joint_angles +- tremors`, the sign flipping with the parity of the
already incremented step counter; :226-229 `enforce_joint_limits` after every substep), executed on the CPU oracle through the
pybullet facade of make_golden_feeding_semantics.py.  A quarter of the benchmarked envs are of this kind.  Output:
tests/golden/feeding_tremor_semantics.npz, replayed by tests/test_reference_feeding_tremor_semantics.py with the repo's restatement
(`tests/parity_cases.apply_tremor` + `take_step_targets` + `feeding_semantics_reference`, what the fused kernels are checked against).

usage: python tests/golden/make_golden_feeding_tremor_semantics.py [/root/reference]"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

N_STEPS, SEED = 12, 9
HEAD = [20, 21, 22, 23]


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    from assistive_gym_b200 import capi
    from assistive_gym_b200.feeding_batch import FeedingBatch
    from oracle.oracle_py import OracleSim
    fb = FeedingBatch()
    sim = OracleSim(fb.scene, capi.default_config(), 1)
    smp = fb.reset(sim, np.random.default_rng(SEED), settle_steps=25, impairment='tremor')
    rest = fb.tremor_rest_of(smp)[0]
    male = bool(smp['male'][0])
    hb = fb.humans['male' if male else 'female']
    from make_golden_env_logic import install_stubs
    from make_golden_feeding_semantics import Facade
    install_stubs(ref)
    import assistive_gym  # noqa: F401  (the reference package)
    from assistive_gym.envs.agents.agent import Agent
    from assistive_gym.envs.feeding_envs import FeedingJacoEnv
    env = FeedingJacoEnv()
    p = sys.modules['pybullet']
    fac = Facade(sim, fb.scene)
    fac.install(p)

    def resetJointState(body, jointIndex=None, targetValue=0.0, targetVelocity=0.0, physicsClientId=None):
        sim.set_joint_state([fac.gl(body, jointIndex)], q=np.array([[float(targetValue)]]), qd=np.array([[float(targetVelocity)]]))
        sim.forward_kinematics()
    p.resetJointState = resetJointState
    env.robot.body, env.tool.body, env.human.body = fb.robot, fb.tool, hb
    env.human.gender = 'male' if male else 'female'
    for a in (env.robot, env.tool, env.human):
        a.id = 0
    sc = fb.scene
    env.robot.controllable_joint_lower_limits = np.array(fb.arm_lower, dtype=np.float64)
    env.robot.controllable_joint_upper_limits = np.array(fb.arm_upper, dtype=np.float64)
    h = env.human
    assert list(h.controllable_joint_indices) == HEAD and not h.controllable
    h.all_joint_indices = list(range(int(sc['body_nlinks'][hb]) - 1))
    h.lower_limits = {j: float(sc['link_lower'][fac.gl(hb, j)]) for j in h.all_joint_indices}
    h.upper_limits = {j: float(sc['link_upper'][fac.gl(hb, j)]) for j in h.all_joint_indices}
    h.controllable_joint_lower_limits = np.array([h.lower_limits[j] for j in HEAD])
    h.controllable_joint_upper_limits = np.array([h.upper_limits[j] for j in HEAD])
    h.impairment, h.tremors, h.strength = 'tremor', np.array(smp['tremors'][0], dtype=np.float64), 1.0
    h.target_joint_angles = np.array(rest, dtype=np.float64)             # human.py:122
    env.robot.motor_gains = env.human.motor_gains = 0.025               # feeding.py:122
    env.agents = [env.robot, env.human]                                 # env.py:130-131
    env.foods = []
    for f in fb.foods:
        a = Agent(); a.body, a.id = f, 0
        env.foods.append(a)
    env.foods_active = list(env.foods)
    env.total_food_count = len(env.foods)
    env.mouth_pos = [0, -0.11, 0.03] if male else [0, -0.1, 0.03]
    env.target = types.SimpleNamespace(set_base_pos_orient=lambda *a, **k: None)
    env.iteration, env.task_success, env.last_sim_time, env.gui = 0, 0, None, False
    env.action_space = types.SimpleNamespace(low=-np.ones(7), high=np.ones(7))
    env.np_random = np.random.RandomState(0)
    env.update_targets()
    start_state = sim.state_get()                                       # stored: the replay does not depend on the IK of the reset being bit-reproducible
    actions = np.random.default_rng(SEED + 1).uniform(-1, 1, size=(N_STEPS, 7)) * 0.3
    obs, rew, head = [], [], []
    for t in range(N_STEPS):
        o, r, d, info = env.step(actions[t].copy())
        obs.append(np.asarray(o, dtype=np.float64)); rew.append(float(r))
        head.append(sim.get_joint_states([fac.gl(hb, j) for j in HEAD])[0][0].copy())
    out = {('sample_' + k): np.asarray(v) for k, v in smp.items()}
    out.update(actions=actions, obs=np.array(obs), reward=np.array(rew), head_q=np.array(head), seed=np.array(SEED))
    out['start_state'] = start_state
    np.savez_compressed(os.path.join(HERE, 'feeding_tremor_semantics.npz'), **out)
    print('tremor amplitudes (deg)', np.round(np.rad2deg(smp['tremors'][0]), 1), 'head - rest (deg) per step, joint 21:', np.round(np.rad2deg(np.array(head)[:, 1] - rest[1]), 2))


if __name__ == '__main__':
    main()
