import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import numpy as np
from assistive_gym_b200 import capi
from assistive_gym_b200.feeding_batch import FeedingBatch
from assistive_gym_b200.sim import BatchSim
lib = capi.load_library(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1] != '-' else capi.load_library()
n = int(os.environ.get("AG_N", "1024"))
fb = FeedingBatch()
h = lambda a: hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:8]
outs = []
for rep in range(6):
    sim = BatchSim(fb.scene, capi.default_config(), n, _lib=lib)
    s = fb.reset(sim, np.random.default_rng(1), settle_steps=0)
    hs = [h(sim.state_get())]
    for i in range(6):
        sim.step(5); hs.append(h(sim.state_get()))
    outs.append(hs); del sim
for hs in outs: print(' '.join(hs))
print('DETERMINISTIC' if all(o == outs[0] for o in outs) else 'NONDETERMINISTIC')
