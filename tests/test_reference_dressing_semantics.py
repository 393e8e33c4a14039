"""`DressingEnv.step` semantics (reference envs/dressing.py:12-106, :199-210 + env.py:174-274 + util.sleeve_on_arm_reward): the
repo's numpy restatement (`tests/dressing_cases.DressingReference`, which the fused Dressing kernels are checked against in
tests/test_dressing.py) replays the rollout of tests/golden/dressing_semantics.npz, produced by the reference's OWN step code on
the CPU oracle (rigid bodies + cloth) through a pybullet facade incl. `getSoftBodyData`
(tests/golden/make_golden_dressing_semantics.py).  Same physics under both; the cloth forces on the person grow to ~12 N."""
import os

import numpy as np

from assistive_gym_b200.dressing_batch import DressingBatch
from oracle.oracle_py import OracleSim
from tests.dressing_cases import DressingReference

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'dressing_semantics.npz'))


def test_restated_dressing_step_reproduces_the_reference_s_rollout():
    db = DressingBatch()
    sim = OracleSim(db.scene, DressingBatch.config(), 1)
    smp = {k[len('sample_'):]: G[k] for k in G.files if k.startswith('sample_')}
    db.reset(sim, np.random.default_rng(0), sample=smp, settle_steps=0)
    sim.cloth_set_gravity([0, 0, -9.81 / 2]); sim.step(3); sim.cloth_set_gravity([0, 0, -9.81])
    ref = DressingReference(db, sim, smp['male'], smp)
    for t, a in enumerate(G['actions']):
        obs, rew, done, info = ref.step(a[None])
        assert np.allclose(obs[0, :23], G['obs'][t][:23], rtol=0, atol=1e-9), (t, np.abs(obs[0] - G['obs'][t]).max())
        assert abs(obs[0, 23] - G['obs'][t][23]) < 1e-9 * (1 + G['obs'][t][23])          # sum of the filtered cloth forces
        assert abs(rew[0] - G['reward'][t]) < 1e-9, (t, rew[0], G['reward'][t])
        assert bool(done[0]) == bool(G['done'][t]) and abs(info[0, 0] - G['total_force'][t]) < 1e-6 * (1 + G['total_force'][t])
        assert int(info[0, 3]) == int(G['sleeve'][t])
    assert G['obs'][:, 23].max() > 5


def test_restated_dressing_step_with_a_tremor_person_reproduces_the_reference_s_rollout():
    """the same with the `tremor` impairment (env.py:130-131, :212-215): the ten left-arm joints shake about their rest targets
    (tests/golden/make_golden_dressing_tremor_semantics.py); `DressingReference`'s tremor branch is what `dressing_pre_body` is checked against"""
    Gt = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'dressing_tremor_semantics.npz'))
    db = DressingBatch()
    sim = OracleSim(db.scene, DressingBatch.config(), 1)
    smp = {k[len('sample_'):]: Gt[k] for k in Gt.files if k.startswith('sample_')}
    assert smp['impairment'][0] == 3 and np.abs(smp['tremors']).max() > 0.1
    db.reset(sim, np.random.default_rng(0), sample=smp, settle_steps=0)
    sim.cloth_set_gravity([0, 0, -9.81 / 2]); sim.step(3); sim.cloth_set_gravity([0, 0, -9.81])
    ref = DressingReference(db, sim, smp['male'], smp)
    for t, a in enumerate(Gt['actions']):
        obs, rew, done, info = ref.step(a[None])
        assert np.allclose(obs[0, :23], Gt['obs'][t][:23], rtol=0, atol=1e-9), (t, np.abs(obs[0] - Gt['obs'][t]).max())
        assert abs(obs[0, 23] - Gt['obs'][t][23]) < 1e-9 * (1 + Gt['obs'][t][23]) and abs(rew[0] - Gt['reward'][t]) < 1e-9
    # the shaking arm changes the rollout: it is not the one without tremor
    assert np.abs(Gt['obs'][:, 14:23] - G['obs'][:, 14:23]).max() > 1e-3


def test_fused_kernel_bodies_reproduce_the_reference_s_rollout(emu_lib):
    """The product's fused Dressing step (`ag_dressing_step_host`: rigid kernels + the cloth step + `dressing_post_body`, compiled for the
    host, fp32) from the golden rollout's start, against what the reference's own `DressingEnv.step` returned on the fp64 oracle -- no
    restatement in between.  The rigid part of the observation agrees to 1e-5; the summed cloth force on the person (nodes crossing
    the 4 cm collision margin one substep earlier or later in fp32) within 35 %, the reward (which carries it at weight 0.01) within 0.02."""
    from assistive_gym_b200.sim import BatchSim
    db = DressingBatch()
    prod = BatchSim(db.scene, DressingBatch.config(), 1, _lib=emu_lib)
    smp = {k[len('sample_'):]: G[k] for k in G.files if k.startswith('sample_')}
    db.reset(prod, np.random.default_rng(0), sample=smp, settle_steps=0)
    prod.cloth_set_gravity([0, 0, -9.81 / 2]); prod.step(3); prod.cloth_set_gravity([0, 0, -9.81])
    db.start_fused(prod, smp)
    for t, a in enumerate(G['actions']):
        obs, rew, done, info = prod.dressing_step_host(a[None].astype(np.float32))
        assert np.abs(obs[0, :23] - G['obs'][t][:23]).max() < 1e-5, (t, np.abs(obs[0, :23] - G['obs'][t][:23]).max())
        assert abs(obs[0, 23] - G['obs'][t][23]) < 0.35 * G['obs'][t][23] + 0.5, (t, obs[0, 23], G['obs'][t][23])
        assert abs(rew[0] - G['reward'][t]) < 0.02 and int(info[0, 3]) == int(G['sleeve'][t])
