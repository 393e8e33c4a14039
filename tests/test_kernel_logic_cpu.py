"""Kernel logic (the per-lane bodies in assistive_gym_b200/csrc/*.cuh, compiled for the host by
tests/kernel_harness) against the CPU oracle.  Runs without a GPU; the same cases run on the real
CUDA build in test_gpu_parity.py."""
import numpy as np
import pytest

from assistive_gym_b200 import capi
from assistive_gym_b200.sim import BatchSim
from tests import parity_cases as pc


@pytest.fixture(scope='module')
def make_sim(emu_lib):
    return lambda scene, cfg, n: BatchSim(scene, cfg, n, _lib=emu_lib)


def test_rollout_200_substeps_strict(feeding, make_sim):
    err = pc.rollout_errors(feeding, make_sim, n=2, seed=0, env_steps=40, foods=False)
    assert err['q'] < pc.TOL_RAD, err
    assert err['tool'] < pc.TOL_M and err['ee'] < pc.TOL_M and err['bowl'] < pc.TOL_M, err


def test_rollout_200_substeps_with_food(feeding, make_sim):
    err = pc.rollout_errors(feeding, make_sim, n=2, seed=1, env_steps=40, foods=True)
    print('foods-on rollout errors', err)
    assert err['q'] < 2e-3 and err['tool'] < 2e-3 and err['ee'] < 2e-3, err


def test_onestep_synchronised(feeding, make_sim):
    err = pc.onestep_errors(feeding, make_sim, n=2, seed=1, steps=20)
    assert err['q'] < 1e-5 and err['tool_pos'] < 1e-5, err
    assert err['pos'] < pc.TOL_M, err


def test_tool_on_body_contact(feeding, make_sim):
    res = pc.tool_contact_case(feeding, make_sim, n=2, seed=2)
    assert res['force'] > 1.0, res            # the case must actually produce tool-on-body contact
    assert res['force_rel'] < pc.TOL_FORCE, res
    assert res['pos'] < pc.TOL_M and res['tool_pos'] < pc.TOL_M, res


def test_fused_feeding_step_semantics(feeding, make_sim):
    """Fused kernels (action -> obs/reward/done) vs the numpy restatement of feeding.py on the oracle."""
    fb = feeding
    n = 2
    cfg = capi.default_config(residual_threshold=0.0)
    cpu, dev, s = pc.synced_pair(fb, make_sim, n, 3, cfg)
    dev.feeding_init(fb.feeding_params(), s['male'])
    st = dict(male=s['male'], foods=np.ones((n, 8), dtype=bool), active=np.ones((n, 8), dtype=bool),
              iteration=np.zeros(n, dtype=int), task_success=np.zeros(n, dtype=int))
    rng = np.random.default_rng(11)
    for k in range(6):
        act = rng.uniform(-1.5, 1.5, size=(n, 7)).astype(np.float32)
        tgt = pc.take_step_targets(cpu.get_joint_states(fb.arm_links)[0], act, fb.arm_lower, fb.arm_upper)
        cpu.set_motor_targets(fb.arm_links, tgt)
        cpu.step(5)
        obs_ref, rew_ref, done_ref, total_ref = pc.feeding_semantics_reference(fb, cpu, act, st)
        obs, rew, done, info = dev.feeding_step_host(act)
        assert np.abs(obs - obs_ref).max() < 1e-3, (k, np.abs(obs - obs_ref).max(axis=0))
        assert np.abs(rew - rew_ref).max() < 2e-3, (k, rew, rew_ref)
        assert np.array_equal(done > 0.5, done_ref)


def test_contact_budget_overflow_flag(feeding, make_sim):
    cfg = capi.default_config(max_contacts=8)
    sim = make_sim(feeding.scene, cfg, 2)
    feeding.reset(sim, np.random.default_rng(0), settle_steps=3)
    assert sim.overflow_count() == 2
