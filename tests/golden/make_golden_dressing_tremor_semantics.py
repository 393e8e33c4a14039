"""(Variant of make_golden_dressing_semantics.py: the person has the `tremor` impairment -- envs/env.py:130-131, :212-215: the ten left-arm
joints are driven to `target_joint_angles +- tremors`, the sign flipping every step; human.py:86-92 draws the amplitudes.)

Golden rollout of the reference's OWN `DressingEnv.step` (envs/dressing.py:12-106, `update_targets` :199-210, env.py:174-274 and
util.sleeve_on_arm_reward, all unmodified) executed on the CPU oracle (rigid bodies + cloth) through the pybullet facade of
make_golden_feeding_semantics.py, extended by `getSoftBodyData` (the fork-only call of dressing.py:25: node positions, contact
positions and contact forces of the cloth, answered from the oracle's cloth).  The reset (robot base-pose search with the
product's IK, host-compiled kernel bodies) is stored.  Output: tests/golden/dressing_tremor_semantics.npz, replayed by
tests/test_reference_dressing_semantics.py with `tests/dressing_cases.DressingReference` (what the fused Dressing kernels are
checked against).

usage: python tests/golden/make_golden_dressing_semantics.py [/root/reference]"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

N_STEPS, SEED = 8, 0


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    from assistive_gym_b200 import capi
    from assistive_gym_b200.dressing_batch import DressingBatch
    from assistive_gym_b200.sim import BatchSim
    from oracle.oracle_py import OracleSim
    db = DressingBatch()
    cfg = DressingBatch.config()
    emu = capi.load_library(os.path.join(ROOT, 'tests', 'kernel_harness', 'libagphys_emu.so'))
    prod = BatchSim(db.scene, cfg, 1, _lib=emu)
    rng = np.random.default_rng(SEED)
    smp = db.sample(1, rng)
    smp['impairment'][:] = 3; smp['strength'] = np.ones(1); smp['tremors'] = np.deg2rad(8.0) * np.sign(rng.uniform(-1, 1, size=(1, 10)))
    smp = db.reset(prod, rng, sample=smp, attempts=12, settle_steps=0)
    assert db.unresolved == 0
    sim = OracleSim(db.scene, cfg, 1)
    db.reset(sim, np.random.default_rng(SEED), sample=smp, settle_steps=0)
    sim.cloth_set_gravity([0, 0, -9.81 / 2]); sim.step(3); sim.cloth_set_gravity([0, 0, -9.81])       # a short settle (dressing.py:178-193)
    male = bool(smp['male'][0])
    from make_golden_env_logic import install_stubs
    from make_golden_feeding_semantics import Facade
    install_stubs(ref)
    import assistive_gym  # noqa: F401  (the reference package)
    from assistive_gym.envs.dressing_envs import DressingPR2Env
    from assistive_gym.envs.util import Util
    env = DressingPR2Env()
    p = sys.modules['pybullet']
    Facade(sim, db.scene, f32_targets=False).install(p)

    def getSoftBodyData(cloth, physicsClientId=None):
        x, _ = sim.cloth_get_state()
        cnt, node, cpos, force, link = sim.cloth_get_contacts(2048)
        k = int(cnt[0])
        return (x[0, :, 0], x[0, :, 1], x[0, :, 2], cpos[0, :k, 0], cpos[0, :k, 1], cpos[0, :k, 2], force[0, :k, 0], force[0, :k, 1], force[0, :k, 2])
    p.getSoftBodyData = getSoftBodyData
    env.robot.body = db.robot
    env.human.body = db.humans['male' if male else 'female']
    env.human.gender = 'male' if male else 'female'
    from assistive_gym_b200.dressing_batch import RADII
    env.human.hand_radius, env.human.elbow_radius, env.human.shoulder_radius = RADII['male' if male else 'female']
    for a in (env.robot, env.human):
        a.id = 0
    env.robot.controllable_joint_lower_limits = np.array(db.arm_lower, dtype=np.float64)
    env.robot.controllable_joint_upper_limits = np.array(db.arm_upper, dtype=np.float64)
    env.robot.motor_gains = env.human.motor_gains = 0.01                 # dressing.py:121
    from assistive_gym_b200.dressing_batch import LEFT_ARM_JOINTS
    sc = db.scene
    h = env.human
    hb = env.human.body
    gl = lambda j: int(sc['body_link0'][hb]) + 1 + j
    h.all_joint_indices = list(range(int(sc['body_nlinks'][hb]) - 1))
    h.lower_limits = {j: float(sc['link_lower'][gl(j)]) for j in h.all_joint_indices}
    h.upper_limits = {j: float(sc['link_upper'][gl(j)]) for j in h.all_joint_indices}
    assert list(h.controllable_joint_indices) == LEFT_ARM_JOINTS and not h.controllable
    h.controllable_joint_lower_limits = np.array([h.lower_limits[j] for j in LEFT_ARM_JOINTS])
    h.controllable_joint_upper_limits = np.array([h.upper_limits[j] for j in LEFT_ARM_JOINTS])
    h.impairment, h.tremors, h.strength = 'tremor', np.array(smp['tremors'][0], dtype=np.float64), 1.0
    h.target_joint_angles = np.array(db.human_rest[0], dtype=np.float64)          # human.py:122
    h.motor_forces = 1.0

    def resetJointState(body, jointIndex=None, targetValue=0.0, targetVelocity=0.0, physicsClientId=None):
        sim.set_joint_state([int(sc['body_link0'][body]) + 1 + int(jointIndex)], q=np.array([[float(targetValue)]]), qd=np.array([[float(targetVelocity)]]))
        sim.forward_kinematics()
    p.resetJointState = resetJointState
    env.agents = [env.robot, env.human]                                  # env.py:130-131
    env.cloth = 0
    env.cloth_attachment = types.SimpleNamespace(set_base_pos_orient=lambda pos, orient: sim.cloth_set_anchor(np.asarray(pos, dtype=np.float64)[None]))
    env.triangle1_point_indices, env.triangle2_point_indices = [1180, 2819, 30], [1322, 13, 696]      # dressing.py:156-157
    env.iteration, env.task_success, env.last_sim_time, env.gui = 0, 0, None, False
    env.action_space = types.SimpleNamespace(low=-np.ones(7), high=np.ones(7))
    env.np_random = np.random.RandomState(0)
    if getattr(env, 'util', None) is None:
        env.util = Util(0, env.np_random)
    arng = np.random.default_rng(SEED + 1)
    actions = arng.uniform(-1, 1, size=(N_STEPS, 7))
    obs, rew, done, total, success, sleeve = [], [], [], [], [], []
    for t in range(N_STEPS):
        o, r, d, info = env.step(actions[t].copy())
        obs.append(np.asarray(o, dtype=np.float64)); rew.append(float(r)); done.append(bool(d)); total.append(float(info['total_force_on_human']))
        success.append(float(env.task_success)); sleeve.append(int(bool(env.forearm_in_sleeve)) + 2 * int(bool(env.upperarm_in_sleeve)))
    out = {('sample_' + k): np.asarray(v) for k, v in smp.items()}
    out.update(actions=actions, obs=np.array(obs), reward=np.array(rew), done=np.array(done), total_force=np.array(total), task_success=np.array(success), sleeve=np.array(sleeve))
    np.savez_compressed(os.path.join(HERE, 'dressing_tremor_semantics.npz'), **out)
    out2 = dict(np.load(os.path.join(HERE, 'dressing_tremor_semantics.npz')))
    print('steps', N_STEPS, 'reward', np.round(rew, 3), 'cloth force sum', np.round([o_[23] for o_ in obs], 2), 'sleeve state', sleeve)


if __name__ == '__main__':
    main()
