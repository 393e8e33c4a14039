"""Golden rollout of the reference's OWN `BedBathingEnv.step` (envs/bed_bathing.py:12-111 + `generate_targets` / `update_targets`
:173-203 + env.py:174-274, unmodified) executed on the CPU oracle through the pybullet facade of make_golden_feeding_semantics.py.
The start state has the wiper pad pressed onto the forearm and the actions keep driving it there, so `get_total_force` wipes
targets.  Output: tests/golden/bathing_semantics.npz, replayed by tests/test_reference_bathing_semantics.py with the repo's
`BedBathingEnv.step_reference_api` + `BedBathingBatch.total_force` (what the fused BedBathing kernels are checked against).

usage: python tests/golden/make_golden_bathing_semantics.py [/root/reference]"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

N_STEPS, SEED = 16, 8


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    from assistive_gym_b200.bed_bathing_batch import SAWYER, BedBathingBatch
    from oracle.oracle_py import OracleSim
    from tests.test_bed_bathing import _pressed_pair
    bb = BedBathingBatch()
    sim, _other, smp, ik = _pressed_pair(bb, lambda sc, cfg, n: OracleSim(sc, cfg, n), 1, seed=SEED)
    start_state = sim.state_get()
    male = bool(smp['male'][0])
    arm = np.array(SAWYER['arm']) + 1
    q_lo = ik[2][:, arm]
    from make_golden_env_logic import install_stubs
    from make_golden_feeding_semantics import Facade
    install_stubs(ref)
    import assistive_gym  # noqa: F401  (the reference package)
    from assistive_gym.envs.bed_bathing_envs import BedBathingSawyerEnv
    env = BedBathingSawyerEnv()
    Facade(sim, bb.scene, f32_targets=True).install(sys.modules['pybullet'])
    env.robot.body, env.tool.body = bb.robot, bb.tool
    env.human.body = bb.humans['male' if male else 'female']
    env.human.gender = 'male' if male else 'female'
    env.human.all_joint_indices = list(range(int(bb.scene['body_nlinks'][env.human.body]) - 1))
    for a in (env.robot, env.tool, env.human):
        a.id = 0
    env.robot.controllable_joint_lower_limits = np.array(bb.arm_lower, dtype=np.float64)
    env.robot.controllable_joint_upper_limits = np.array(bb.arm_upper, dtype=np.float64)
    env.robot.motor_gains, env.robot.motor_forces = 0.1, 5.0             # a stronger arm than robot.py:36-37 so that the pad reaches the skin within the rollout
    env.agents = [env.robot]
    env.iteration, env.task_success, env.last_sim_time, env.gui = 0, 0, None, False
    env.action_space = types.SimpleNamespace(low=-np.ones(7), high=np.ones(7))
    env.np_random = np.random.RandomState(0)
    if getattr(env, 'util', None) is None:
        from assistive_gym.envs.util import Util
        env.util = Util(0, env.np_random)
    env.create_spheres = lambda radius=0.01, mass=0.0, batch_positions=(), **k: [types.SimpleNamespace(set_base_pos_orient=lambda *a, **kk: None) for _ in batch_positions]
    env.generate_targets()
    actions, obs, rew, done, total, on_human, new_pts, success = [], [], [], [], [], [], [], []
    for t in range(N_STEPS):
        q = sim.get_joint_states(bb.arm_links)[0]
        a = np.clip((q_lo - q) / 0.25, -1, 1)[0]                        # keep pressing (tests/test_bed_bathing._check_fused_wiping)
        o, r, d, info = env.step(a.copy())
        actions.append(a); obs.append(np.asarray(o, dtype=np.float64)); rew.append(float(r)); done.append(bool(d)); total.append(float(info['total_force_on_human']))
        on_human.append(float(env.tool_force_on_human)); new_pts.append(int(env.new_contact_points)); success.append(int(env.task_success))
    out = {('sample_' + k): np.asarray(v) for k, v in smp.items()}
    out.update(start_state=start_state, q_press=q_lo, actions=np.array(actions), obs=np.array(obs), reward=np.array(rew), done=np.array(done), total_force=np.array(total),
               tool_force_on_human=np.array(on_human), new_contact_points=np.array(new_pts), task_success=np.array(success), total_target_count=np.array(env.total_target_count), motor_gain=np.array(0.1), motor_force=np.array(5.0))
    np.savez_compressed(os.path.join(HERE, 'bathing_semantics.npz'), **out)
    print('steps', N_STEPS, 'targets', env.total_target_count, 'wiped per step', new_pts, 'cloth force', np.round(on_human, 2), 'reward', np.round(rew, 2))


if __name__ == '__main__':
    main()
