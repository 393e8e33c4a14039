import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import numpy as np
from assistive_gym_b200 import capi
from assistive_gym_b200.feeding_batch import FeedingBatch
from assistive_gym_b200.sim import BatchSim
lib = capi.load_library(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1] != '-' else capi.load_library()
n = int(os.environ.get('AG_N', '4'))
fb = FeedingBatch()
sim = BatchSim(fb.scene, capi.default_config(), n, _lib=lib)
fb.reset(sim, np.random.default_rng(11), settle_steps=5)
st = sim.state_get()
sim.step(5); a = sim.state_get()
nb = fb.scene.n_bodies
bad = 0
for k in range(30):
    sim.state_set(st); sim.step(5); b = sim.state_get()
    if not np.array_equal(a, b):
        bad += 1
        d = np.abs(a - b); e, c = np.unravel_index(np.argmax(d), d.shape)
        what = ('body %d comp %d' % (c // 13, c % 13)) if c < nb * 13 else ('link %d %s' % ((c - nb * 13) // 2, 'q' if (c - nb * 13) % 2 == 0 else 'qd'))
        cnt, it = sim.solver_stats(); t, f = sim.pgs_trips()
        print('rep', k, 'max diff %.3g env %d %s; envs differing %s; iters %s floats %s' % (d.max(), e, what, np.nonzero(d.max(axis=1) > 0)[0], it, f))
print('bad', bad, 'of 30')
