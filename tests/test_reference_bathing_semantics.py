"""`BedBathingEnv.step` semantics (reference envs/bed_bathing.py:12-111, :173-203 + env.py:174-274): the repo's per-call restatement
(`BedBathingEnv.step_reference_api` + `BedBathingBatch.total_force` / `targets_world`, what the fused BedBathing kernels are checked
against in tests/test_bed_bathing.py) replays the rollout of tests/golden/bathing_semantics.npz, produced by the reference's OWN
step code on the CPU oracle through a pybullet facade (tests/golden/make_golden_bathing_semantics.py): the wiper pad is pressed
onto the forearm, two wiping targets are cleared, the cloth force ramps to 6 N.  Same physics under both (the oracle)."""
import os

import numpy as np

from assistive_gym_b200 import envs
from assistive_gym_b200.bed_bathing_batch import BedBathingBatch
from assistive_gym_b200.envs.agents.furniture import Furniture
from oracle.oracle_py import OracleSim
from tests.test_bed_bathing import _pressed_pair

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'bathing_semantics.npz'))


def test_restated_bed_bathing_step_reproduces_the_reference_s_rollout():
    bb = BedBathingBatch()
    sim, _other, smp, ik = _pressed_pair(bb, lambda sc, cfg, n: OracleSim(sc, cfg, n), 1, seed=8)
    assert np.allclose(sim.state_get(), G['start_state'], atol=1e-9)            # the generator's start state ...
    sim.state_set(G['start_state']); sim.forward_kinematics()                  # ... to the last bit (the IK that builds it goes through BLAS)
    env = envs.make('BedBathingSawyer-v1', n_envs=1)
    env._bb = bb
    env.id = sim                                                               # the env's per-call path on the oracle instead of the CUDA library
    env.plane.init(bb.plane, sim, env.np_random, indices=-1)
    env.robot.init(bb.robot, sim, env.np_random)
    env.tool.init(bb.tool, sim, env.np_random, indices=-1)
    env.furniture.init(bb.bed, sim, env.np_random, indices=-1)
    env.humans = {}
    for g, hb in bb.humans.items():
        h = type(env.human)(env.human.controllable_joint_indices, controllable=False)
        h.init(hb, sim, env.np_random, env.human.controllable_joint_indices)
        env.humans[g] = h
    env.agents = [env.robot]
    env.robot.motor_gains, env.robot.motor_forces = float(G['motor_gain']), float(G['motor_force'])
    env.male = smp['male'].astype(bool)
    env.targets_pos_world, env.targets_alive = bb.targets_world(sim, smp)
    env.total_target_count = env.targets_alive.sum(axis=1)
    assert int(env.total_target_count[0]) == int(G['total_target_count'])
    env.task_success = np.zeros(1, dtype=int)
    env.iteration = 0
    for t, a in enumerate(G['actions']):
        obs, rew, done, info = env.step_reference_api(a[None])
        # (the host mirror hands the motor targets over as fp32, the device's type: the two rollouts differ at the 1e-8 level)
        assert np.allclose(obs[:23], G['obs'][t][:23], rtol=0, atol=1e-6), (t, np.abs(obs - G['obs'][t]).max())
        # forces: the reference sums the fp32 contact records, the restatement asks the oracle for the fp64 sum
        assert abs(obs[23] - G['obs'][t][23]) < 1e-4 * (1 + abs(G['obs'][t][23]))
        assert abs(rew - G['reward'][t]) < 1e-4, (t, rew, G['reward'][t])
        assert bool(done) == bool(G['done'][t]) and abs(info['total_force_on_human'] - G['total_force'][t]) < 1e-4 * (1 + G['total_force'][t])
        assert abs(env.tool_force_on_human[0] - G['tool_force_on_human'][t]) < 1e-4 * (1 + G['tool_force_on_human'][t])
        assert int(env.new_contact_points[0]) == int(G['new_contact_points'][t]) and int(env.task_success[0]) == int(G['task_success'][t])
    assert G['new_contact_points'].sum() >= 2 and G['tool_force_on_human'].max() > 3


def test_fused_kernel_bodies_reproduce_the_reference_s_rollout(emu_lib):
    """The product's fused BedBathing step (`ag_bathing_step_host`, kernel bodies compiled for the host, fp32) from the golden rollout's
    start, against what the reference's own `BedBathingEnv.step` returned on the fp64 oracle -- no restatement in between: the pad is
    pressed onto the forearm, the same two targets are wiped in the same step, forces within 5 %."""
    from assistive_gym_b200.sim import BatchSim
    bb = BedBathingBatch()
    cpu, prod, smp, ik = _pressed_pair(bb, lambda sc, cfg, n: BatchSim(sc, cfg, n, _lib=emu_lib), 1, seed=8)
    prod.state_set(G['start_state'].astype(np.float32)); prod.forward_kinematics()
    prod.set_motor(bb.arm_links, 1, target=prod.get_joint_states(bb.arm_links)[0], kp=[float(G['motor_gain'])] * 7, kd=[1.0] * 7, max_force=[float(G['motor_force'])] * 7)
    tw, alive = bb.targets_world(cpu, smp)
    prod.bathing_init(bb.bathing_params(), smp['male'], tw, alive)
    wiped = 0
    for t, a in enumerate(G['actions']):
        obs, rew, done, info = prod.bathing_step_host(a[None].astype(np.float32))
        assert np.abs(obs[0, :23] - G['obs'][t][:23]).max() < 2e-3, (t, np.abs(obs[0, :23] - G['obs'][t][:23]).max())
        assert abs(info[0, 2] - G['tool_force_on_human'][t]) < 0.05 * G['tool_force_on_human'][t] + 0.05, (t, info[0, 2], G['tool_force_on_human'][t])
        assert int(info[0, 3]) == int(G['new_contact_points'][t]) and abs(rew[0] - G['reward'][t]) < 0.05
        wiped += int(info[0, 3])
    assert wiped == int(G['new_contact_points'].sum()) >= 2
