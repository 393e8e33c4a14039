#!/usr/bin/env python3
"""Diagnostic (GPU box): run-to-run determinism of the bench workload: hash of the state after reset and after each step."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from assistive_gym_b200 import capi
from assistive_gym_b200.feeding_batch import FeedingBatch
from assistive_gym_b200.sim import BatchSim
from assistive_gym_b200.sharding import sample_block
n = int(os.environ.get('AG_N', '4096'))
fb = FeedingBatch(); sim = BatchSim(fb.scene, capi.default_config(), n)
rng = np.random.default_rng(1001)
s = fb.reset(sim, rng, settle_steps=0, sample=sample_block(fb, 0, n))
h = lambda a: hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:10]
print('reset', h(sim.state_get()))
sim.step(25)
print('settle', h(sim.state_get()))
fb.start_fused(sim, s, seed=1001)
arng = np.random.default_rng(0)
prev = sim.state_get()
for i in range(int(os.environ.get('AG_STEPS', '13'))):
    obs, rew, done, info = sim.feeding_step_host(arng.uniform(-1, 1, size=(n, 7)).astype(np.float32))
    st = sim.state_get(); cnt, it = sim.solver_stats()
    print('step', i, h(st), h(obs), 'contacts', int(cnt.sum()), 'iters', int(it.sum()), 'overflow', sim.overflow_count())
    np.save(os.path.join(ROOT, 'gpurun_out', 'det_%s_%d.npy' % (os.environ.get('AG_TAG', 'a'), i)), st)
