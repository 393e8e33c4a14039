"""GPU parity tests proper: the CUDA build (through the C ABI) against the CPU oracle, plus
size-independent properties at the benchmark batch size.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

from assistive_gym_b200 import capi
from assistive_gym_b200.sim import BatchSim
from tests import parity_cases as pc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def make_sim(gpu_lib):
    return lambda scene, cfg, n: BatchSim(scene, cfg, n, device=0)


def test_rollout_200_substeps_strict(feeding, make_sim):
    err = pc.rollout_errors(feeding, make_sim, n=8, seed=0, env_steps=40, foods=False)
    print('strict rollout errors', err)
    assert err['q'] < pc.TOL_RAD, err
    assert err['tool'] < pc.TOL_M and err['ee'] < pc.TOL_M and err['bowl'] < pc.TOL_M, err


def test_rollout_200_substeps_with_food(feeding, make_sim):
    err = pc.rollout_errors(feeding, make_sim, n=8, seed=1, env_steps=40, foods=True)
    print('foods-on rollout errors', err)
    assert err['q'] < 2e-3 and err['tool'] < 2e-3 and err['ee'] < 2e-3, err


def test_rollout_tremor_head(feeding, make_sim):
    """impairment == tremor in every env: simulated 4-DoF head chain, +-tremor targets, hard limits."""
    err = pc.rollout_errors(feeding, make_sim, n=8, seed=6, env_steps=40, foods=False, impairment='tremor')
    print('tremor rollout errors', err)
    assert err['head_travel'] > 0.02, err
    assert err['head'] < pc.TOL_RAD and err['q'] < pc.TOL_RAD, err
    assert err['tool'] < pc.TOL_M and err['bowl'] < pc.TOL_M, err


def test_rollout_strict_population(feeding, make_sim):
    """64 envs x 100 substeps.  PGS stops after exactly 50 iterations; in rare steps an active-set
    switch (a contact opening/closing) lands on the last iteration in fp64 but not in fp32 and the
    arm takes a ~1e-4..1e-3 rad kick (both converge to the same answer with more iterations).  The
    strict tolerance is therefore asserted on the population: the median env and >= 90 % of envs."""
    err = pc.rollout_errors(feeding, make_sim, n=64, seed=9, env_steps=20, foods=False)
    q = err['q_env']
    print('population: median %.3g  p90 %.3g  max %.3g  within tol %.3f' % (np.median(q), np.quantile(q, 0.9), q.max(), (q < pc.TOL_RAD).mean()))
    assert np.median(q) < 1e-5 and (q < pc.TOL_RAD).mean() >= 0.9 and q.max() < 5e-3, err


def test_benchmarked_config_population(feeding, make_sim):
    """parity ON THE CONFIGURATION bench.py measures (foods on, early exit 1e-7, random actions)
    at n = 1024 over 200 substeps, reported as a distribution, with the oracle's own fp32 build as the control.
    Measured on the B200: the free-running rollout of this configuration is chaotic at the level
    of the north-star tolerance -- the fp32 build of the ORACLE ITSELF ends up a median 1.4e-3 rad from its fp64 build --
    so what can be asserted is (a) the CUDA build is as close to the fp64 oracle as the oracle's fp32 build is, quantile
    by quantile, and (b) re-synchronised every env step, the CUDA build meets the tolerances for (nearly) every env."""
    import json
    import os
    thr = max(1, len(os.sched_getaffinity(0)))
    e = pc.population_errors(feeding, make_sim, n=1024, seed=21, env_steps=40, threads=thr)
    prod, ctrl = pc.population_summary(e['product']), pc.population_summary(e['oracle_f32'])
    print('population n=1024 foods on early exit: product', json.dumps(prod))
    print('population n=1024 foods on early exit: oracle fp32 control', json.dumps(ctrl))
    for k in ('q', 'tool', 'ee'):
        assert prod[k]['within'] >= ctrl[k]['within'] - 0.05, (k, prod[k], ctrl[k])
        for qn in ('median', 'p90', 'p99'):
            assert prod[k][qn] <= 1.5 * ctrl[k][qn] + 1e-6, (k, qn, prod[k], ctrl[k])


def test_benchmarked_config_resynchronised_population(feeding, make_sim):
    """The same configuration at n = 1024, state copied from the oracle before every env step (5 substeps with the
    default early exit): the step function itself, without chaotic drift."""
    import json
    import os
    n = 1024
    cfg = capi.default_config()
    cpu, dev, s = pc.synced_pair(feeding, make_sim, n, 31, cfg, threads=max(1, len(os.sched_getaffinity(0))))
    fb = feeding
    L = pc.feeding_links(fb)
    links = [L['tool'], L['ee']]
    rng = np.random.default_rng(5)
    eq, et = np.zeros(n), np.zeros(n)
    for it in range(8):
        act = rng.uniform(-1, 1, size=(n, 7))
        tgt = pc.take_step_targets(cpu.get_joint_states(fb.arm_links)[0], act, fb.arm_lower, fb.arm_upper)
        dev.state_set(cpu.state_get())
        for sim in (cpu, dev):
            sim.set_motor_targets(fb.arm_links, tgt)
            sim.step(5)
        eq = np.maximum(eq, np.abs(cpu.get_joint_states(fb.arm_links)[0] - dev.get_joint_states(fb.arm_links)[0]).max(axis=1))
        et = np.maximum(et, np.abs(cpu.get_link_states(links)['pos'] - dev.get_link_states(links)['pos']).max(axis=(1, 2)))
    res = dict(q=dict(median=float(np.median(eq)), p99=float(np.quantile(eq, 0.99)), max=float(eq.max()), within=float((eq < pc.TOL_RAD).mean())),
               pos=dict(median=float(np.median(et)), p99=float(np.quantile(et, 0.99)), max=float(et.max()), within=float((et < pc.TOL_M).mean())))
    print('resynchronised population n=1024 foods on early exit:', json.dumps(res))
    assert res['q']['median'] < 1e-5 and res['pos']['median'] < 1e-5, res
    assert res['q']['within'] >= 0.97 and res['pos']['within'] >= 0.99, res


def test_onestep_synchronised(feeding, make_sim):
    err = pc.onestep_errors(feeding, make_sim, n=8, seed=1, steps=30)
    print('one-step errors', err)
    assert err['q'] < 1e-5 and err['tool_pos'] < 1e-5, err
    assert err['pos'] < pc.TOL_M, err


def test_onestep_default_early_exit(feeding, make_sim):
    """Default residual threshold (1e-7): the early-exit decision is discontinuous, so only the
    well-conditioned quantities are bounded tightly."""
    err = pc.onestep_errors(feeding, make_sim, n=8, seed=4, steps=20, residual_threshold=1e-7)
    print('one-step errors (early exit)', err)
    assert err['q'] < 1e-3 and err['tool_pos'] < pc.TOL_M, err


def test_tool_on_body_contact(feeding, make_sim):
    res = pc.tool_contact_case(feeding, make_sim, n=8, seed=2)
    print('tool contact', res)
    assert res['force'] > 1.0, res
    assert res['force_rel'] < pc.TOL_FORCE, res
    assert res['pos'] < pc.TOL_M and res['tool_pos'] < pc.TOL_M, res


@pytest.mark.parametrize('impairment', ['random', 'tremor'])
def test_fused_feeding_step_semantics(feeding, make_sim, impairment):
    fb = feeding
    n = 8
    cfg = capi.default_config(residual_threshold=0.0)
    cpu, dev, s = pc.synced_pair(fb, make_sim, n, 3, cfg, impairment=impairment)
    fb.start_fused(dev, s)
    st = dict(male=s['male'], foods=np.ones((n, 8), dtype=bool), active=np.ones((n, 8), dtype=bool),
              iteration=np.zeros(n, dtype=int), task_success=np.zeros(n, dtype=int))
    rng = np.random.default_rng(11)
    for k in range(6):
        act = rng.uniform(-1.5, 1.5, size=(n, 7)).astype(np.float32)
        tgt = pc.take_step_targets(cpu.get_joint_states(fb.arm_links)[0], act, fb.arm_lower, fb.arm_upper)
        cpu.set_motor_targets(fb.arm_links, tgt)
        pc.apply_tremor(fb, (cpu,), s, k + 1)
        cpu.step(5)
        obs_ref, rew_ref, done_ref, _ = pc.feeding_semantics_reference(fb, cpu, act, st)
        obs, rew, done, info = dev.feeding_step_host(act)
        # foods are on and the actions are large: a 1 g sphere bouncing differently in fp32 and fp64 can flip a contact
        # (and a +20 / -5 food event) in a single env, so at most one env of the eight may leave the tight bounds
        eo, er = np.abs(obs - obs_ref).max(axis=1), np.abs(rew - rew_ref)
        assert (eo >= 1e-3).sum() <= 1 and np.median(eo) < 1e-4, (k, eo)
        assert (er >= 2e-3).sum() <= 1 and np.median(er) < 1e-3, (k, er)
        assert np.array_equal(done > 0.5, done_ref)


def test_readback_calls(feeding, make_sim):
    """A4-A7: getJointStates / getLinkState / getContactPoints / getClosestPoints equivalents vs the oracle."""
    r = pc.readback_errors(feeding, make_sim, n=8)
    print('read-back errors', r)
    assert r['count_equal'] and r['n_contacts'] > 0, r
    assert r['q'] < 1e-6 and r['qd'] < 1e-4 and r['tau'] < 1e-3 * max(1.0, r['tau_max']), r
    assert r['pos'] < 1e-5 and r['com_pos'] < 1e-5 and r['quat'] < 1e-5 and r['com_quat'] < 1e-5, r
    assert r['lin_vel'] < 1e-3 and r['ang_vel'] < 1e-3, r
    assert r['contact_pos'] < 1e-4 and r['contact_force'] < 0.05 * 9.81, r
    assert r['closest_dist'] < 1e-5, r


def test_golden_fixture(feeding, make_sim):
    """Committed oracle-generated fixture (tests/golden/make_golden.py): state after 10 substeps."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'feeding_10substeps.npz'))
    fb = feeding
    n = z['state0'].shape[0]
    cfg = capi.default_config(residual_threshold=0.0)
    dev = make_sim(fb.scene, cfg, n)
    sample = {k[2:]: z[k] for k in z.files if k.startswith('s_')}
    fb.reset(dev, np.random.default_rng(0), settle_steps=0, sample=sample)
    dev.state_set(z['state0'])
    dev.set_motor_targets(fb.arm_links, z['targets'])
    dev.step(10)
    st = dev.state_get().astype(np.float64)
    nb = fb.scene.n_bodies
    dq = np.abs(st[:, nb * 13:] - z['state10'][:, nb * 13:]).reshape(n, -1, 2)[:, :, 0].max()
    dtool = np.abs(st[:, fb.tool * 13:fb.tool * 13 + 3] - z['state10'][:, fb.tool * 13:fb.tool * 13 + 3]).max()
    assert dq < pc.TOL_RAD and dtool < pc.TOL_M, (dq, dtool)


def test_batch4096_properties(feeding, make_sim):
    """BASELINE.json batch size: replicated envs stay bit-identical (lock-step determinism), different
    envs stay finite and normalised, the contact budget is not exceeded, foods stay in the spoon at rest."""
    fb = feeding
    n = 4096
    cfg = capi.default_config()
    dev = make_sim(fb.scene, cfg, n)
    rng = np.random.default_rng(5)
    s = fb.sample(n, rng)
    for k in s:                 # envs [0:64) replicated into [64:128)
        s[k][64:128] = s[k][0:64]
    import copy
    rep = np.random.default_rng(9)
    fb.reset(dev, rep, settle_steps=0, sample=s)
    st = dev.state_get()
    st[64:128] = st[0:64]
    dev.state_set(st)
    q0 = dev.get_joint_states(fb.arm_links)[0]
    dev.set_motor_targets(fb.arm_links, q0)
    dev.step(25)
    fb.start_fused(dev, s)
    act = np.zeros((n, 7), dtype=np.float32)
    arng = np.random.default_rng(1)
    for i in range(4):
        act = arng.uniform(-1, 1, size=(n, 7)).astype(np.float32)
        act[64:128] = act[0:64]
        obs, rew, done, info = dev.feeding_step_host(act)
    st = dev.state_get()
    assert np.all(np.isfinite(st)) and np.all(np.isfinite(obs)) and np.all(np.isfinite(rew))
    assert np.array_equal(st[0:64], st[64:128])            # bit-exact replication => deterministic ordering
    assert np.array_equal(obs[0:64], obs[64:128])
    nb = fb.scene.n_bodies
    quat = st[:, :nb * 13].reshape(n, nb, 13)[:, :, 3:7]
    assert np.abs(np.linalg.norm(quat, axis=-1) - 1).max() < 1e-4
    # sticky flags: a handful of envs whose start pose is still in collision (IK resampling exhausted) exceed the
    # 128-contact budget while the food settles; their contacts are truncated by key (deterministic)
    assert dev.overflow_count() <= n // 500
    assert obs.shape == (n, 25)
