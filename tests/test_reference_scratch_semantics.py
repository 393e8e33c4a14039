"""`ScratchItchEnv.step` semantics (reference envs/scratch_itch.py:10-91 + env.py:174-274): the repo's numpy restatement
(`tests/test_scratch_itch.ScratchReference`, which the fused ScratchItch kernels are checked against) replays the rollout of
tests/golden/scratch_semantics.npz, produced by the reference's OWN step code on the CPU oracle through a pybullet facade
(tests/golden/make_golden_scratch_semantics.py).  The start state has the tool tip on the itch target, so the scratch
bookkeeping fires; same physics under both, so everything agrees to rounding."""
import os

import numpy as np

from assistive_gym_b200 import capi
from assistive_gym_b200.scratch_itch_batch import ScratchItchBatch
from oracle.oracle_py import OracleSim
from tests.test_scratch_itch import ScratchReference

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'scratch_semantics.npz'))


def test_restated_scratch_itch_step_reproduces_the_reference_s_rollout():
    sb = ScratchItchBatch()
    sim = OracleSim(sb.scene, capi.default_config(residual_threshold=0.0), 1)
    smp = {k[len('sample_'):]: G[k] for k in G.files if k.startswith('sample_')}
    sb.reset(sim, np.random.default_rng(0), sample=smp)
    ref = ScratchReference(sb, sim, smp)
    for t, a in enumerate(G['actions']):
        obs, rew, done, info = ref.step(a[None])
        assert np.allclose(obs[0, :29], G['obs'][t][:29], rtol=0, atol=1e-9), (t, np.abs(obs[0] - G['obs'][t]).max())
        # forces: the reference sums the per-contact forces of the fp32 contact records, the restatement asks the oracle for the fp64 sum
        assert abs(obs[0, 29] - G['obs'][t][29]) < 1e-6 * (1 + abs(G['obs'][t][29]))
        assert abs(rew[0] - G['reward'][t]) < 1e-6, (t, rew[0], G['reward'][t])
        assert bool(done[0]) == bool(G['done'][t])
        assert abs(info[0, 0] - G['total_force'][t]) < 1e-6 * (1 + G['total_force'][t]) and abs(info[0, 2] - G['force_at_target'][t]) < 1e-6 * (1 + G['force_at_target'][t])
        assert int(info[0, 3]) == int(G['task_success'][t])
    assert G['task_success'][-1] >= 1 and G['reward'].max() > 4                 # the rollout contains counted scratches


def test_fused_kernel_bodies_reproduce_the_reference_s_rollout(emu_lib):
    """The product's fused ScratchItch step (`ag_scratch_step_host`: the CUDA kernel bodies compiled for the host, fp32) from the golden
    rollout's start, against what the reference's own `ScratchItchEnv.step` returned on the fp64 oracle -- no restatement in between:
    observation 1e-5, tool force and reward 1 %, the same scratches counted."""
    from assistive_gym_b200.sim import BatchSim
    sb = ScratchItchBatch()
    prod = BatchSim(sb.scene, capi.default_config(residual_threshold=0.0), 1, _lib=emu_lib)
    smp = {k[len('sample_'):]: G[k] for k in G.files if k.startswith('sample_')}
    sb.reset(prod, np.random.default_rng(0), sample=smp)
    sb.start_fused(prod, smp)
    for t, a in enumerate(G['actions']):
        obs, rew, done, info = prod.scratch_step_host(a[None].astype(np.float32))
        assert np.abs(obs[0, :29] - G['obs'][t][:29]).max() < 1e-5, (t, np.abs(obs[0, :29] - G['obs'][t][:29]).max())
        assert abs(obs[0, 29] - G['obs'][t][29]) < 0.01 * abs(G['obs'][t][29]) + 1e-3
        assert abs(rew[0] - G['reward'][t]) < 1e-3 and int(info[0, 3]) == int(G['task_success'][t])
