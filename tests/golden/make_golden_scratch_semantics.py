"""Golden rollout of the reference's OWN `ScratchItchEnv.step` (envs/scratch_itch.py:10-91 + env.py:174-274, unmodified) executed on
the CPU oracle through the pybullet facade of make_golden_feeding_semantics.py; the start state has the tool tip pressed onto the
itch target so that `get_total_force` (force at the target, scratch bookkeeping) is live.  The reset that produces it needs the
product's IK (host-compiled kernel bodies); its outcome is stored.  Output: tests/golden/scratch_semantics.npz, replayed by
tests/test_reference_scratch_semantics.py with the repo's restatement (`tests/test_scratch_itch.ScratchReference`, the function the
fused ScratchItch kernels are checked against).

usage: python tests/golden/make_golden_scratch_semantics.py [/root/reference]"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

N_STEPS, SEED = 16, 3


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    from assistive_gym_b200 import capi
    from assistive_gym_b200.scratch_itch_batch import ScratchItchBatch
    from assistive_gym_b200.sim import BatchSim
    from oracle.oracle_py import OracleSim
    from tests.test_scratch_itch import _press_tip_on_target
    sb = ScratchItchBatch()
    cfg = capi.default_config(residual_threshold=0.0)
    emu = capi.load_library(os.path.join(ROOT, 'tests', 'kernel_harness', 'libagphys_emu.so'))
    prod = BatchSim(sb.scene, cfg, 1, _lib=emu)
    smp = None
    for seed in range(SEED, SEED + 40):                                 # a start pose from which the IK can press the tip onto the target
        smp = sb.reset(prod, np.random.default_rng(seed))
        q7, ok = _press_tip_on_target(sb, prod, smp)
        if ok[0]:
            smp['q7'] = q7
            break
    assert ok[0]
    sim = OracleSim(sb.scene, cfg, 1)
    sb.reset(sim, np.random.default_rng(0), sample=smp)
    male = bool(smp['male'][0])
    from make_golden_env_logic import install_stubs
    from make_golden_feeding_semantics import Facade
    install_stubs(ref)
    import assistive_gym  # noqa: F401  (the reference package)
    from assistive_gym.envs.scratch_itch_envs import ScratchItchJacoEnv
    env = ScratchItchJacoEnv()
    Facade(sim, sb.scene).install(sys.modules['pybullet'])
    env.robot.body, env.tool.body = sb.robot, sb.tool
    env.human.body = sb.humans['male' if male else 'female']
    env.human.gender = 'male' if male else 'female'
    for a in (env.robot, env.tool, env.human):
        a.id = 0
    env.robot.controllable_joint_lower_limits = np.array(sb.arm_lower, dtype=np.float64)
    env.robot.controllable_joint_upper_limits = np.array(sb.arm_upper, dtype=np.float64)
    env.agents = [env.robot]
    env.limb = int(smp['limb_joint'][0])
    env.target_on_arm = np.array(smp['target_local'][0], dtype=np.float64)
    env.target = types.SimpleNamespace(set_base_pos_orient=lambda *a, **k: None)
    env.prev_target_contact_pos = np.zeros(3)
    env.iteration, env.task_success, env.last_sim_time, env.gui = 0, 0, None, False
    env.action_space = types.SimpleNamespace(low=-np.ones(7), high=np.ones(7))
    env.update_targets()
    arng = np.random.default_rng(SEED + 1)
    actions = np.concatenate([np.zeros((4, 7)), arng.uniform(-1, 1, size=(N_STEPS - 4, 7)) * 0.15])      # hold the press, then wiggle
    obs, rew, done, total, at_target, success = [], [], [], [], [], []
    for t in range(N_STEPS):
        o, r, d, info = env.step(actions[t].copy())
        obs.append(np.asarray(o, dtype=np.float64)); rew.append(float(r)); done.append(bool(d)); total.append(float(info['total_force_on_human']))
        at_target.append(float(env.tool_force_at_target)); success.append(int(env.task_success))
    out = {('sample_' + k): np.asarray(v) for k, v in smp.items()}
    out.update(actions=actions, obs=np.array(obs), reward=np.array(rew), done=np.array(done), total_force=np.array(total), force_at_target=np.array(at_target),
               task_success=np.array(success))
    np.savez_compressed(os.path.join(HERE, 'scratch_semantics.npz'), **out)
    print('steps', N_STEPS, 'scratches counted', success[-1], 'force at target', np.round(at_target, 2), 'reward', np.round(rew, 2))


if __name__ == '__main__':
    main()
