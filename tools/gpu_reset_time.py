#!/usr/bin/env python3
"""GPU box: wall-clock of one batched reset next to the 200 steps of the episode it starts."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from assistive_gym_b200 import capi
from assistive_gym_b200.feeding_batch import FeedingBatch
from assistive_gym_b200.sim import BatchSim
n = int(os.environ.get('AG_N', '4096'))
fb = FeedingBatch(); sim = BatchSim(fb.scene, capi.default_config(), n)
rng = np.random.default_rng(0)
for rep in range(2):
    t0 = time.time(); s = fb.reset(sim, rng, settle_steps=25); fb.start_fused(sim, s); t1 = time.time()
    print('reset %d: %.2f s (IK resamples %d, still colliding %d, max IK err %.4f)' % (rep, t1 - t0, fb.ik_resamples, fb.ik_colliding, fb.ik_err.max()))
t0 = time.time()
for i in range(20): sim.feeding_step_host(rng.uniform(-1, 1, size=(n, 7)).astype(np.float32))
print('20 steps through the host API: %.2f s -> 200 steps = %.1f s' % (time.time() - t0, (time.time() - t0) * 10))
