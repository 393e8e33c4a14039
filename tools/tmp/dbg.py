import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import numpy as np
from assistive_gym_b200 import capi
from assistive_gym_b200.feeding_batch import FeedingBatch
from assistive_gym_b200.sim import BatchSim
lib = capi.load_library(os.path.join(ROOT, 'tools', 'tmp', 'libagphys_dbg.so'))
fb = FeedingBatch(); n = 64
sim = BatchSim(fb.scene, capi.default_config(), n, _lib=lib)
s = fb.reset(sim, np.random.default_rng(0), settle_steps=3)
for i in range(4):
    sim.step(1)
    t, f = sim.pgs_trips()
    print('step', i, 'bad records', (t >> 16)[:16], 'first bad trip', (t & 0xffff)[:16], 'floats', f[:8])
