"""Workload of tests/test_sanitizer.py: runs inside a subprocess with libasan preloaded and the kernel bodies compiled
with -fsanitize=address (every "device" buffer is a host calloc there, so an out-of-bounds index in a kernel body or in
the C-ABI glue aborts the process)."""
import sys

import numpy as np

sys.path.insert(0, sys.argv[2])
from assistive_gym_b200 import capi                                   # noqa: E402
from assistive_gym_b200.bed_bathing_batch import BedBathingBatch      # noqa: E402
from assistive_gym_b200.feeding_batch import FeedingBatch             # noqa: E402
from assistive_gym_b200.sim import BatchSim                           # noqa: E402

lib = capi.load_library(sys.argv[1])
fb = FeedingBatch()
for n in (1, 3):                                                      # odd / tiny batches: ragged tails of every grid
    sim = BatchSim(fb.scene, capi.default_config(), n, _lib=lib)
    rng = np.random.default_rng(n)
    s = fb.reset(sim, rng, settle_steps=6)
    fb.start_fused(sim, s)
    for _ in range(2):
        sim.feeding_step_host(rng.uniform(-1, 1, size=(n, 7)).astype(np.float32))
    sim.get_contacts(fb.tool, max_pts=8)
    sim.closest_points(fb.robot, fb.table, 0.5, max_pts=4)
    sim.state_set(sim.state_get())
    sim.close()
sim = BatchSim(fb.scene, capi.default_config(max_contacts=8), 3, _lib=lib)      # over-budget paths
fb.reset(sim, np.random.default_rng(0), settle_steps=4)
assert sim.overflow_count() == 3
sim.close()
bb = BedBathingBatch()
sim = BatchSim(bb.scene, capi.default_config(), 3, _lib=lib)
s = bb.reset(sim, np.random.default_rng(1))
bb.start_fused(sim, s)
sim.bathing_step_host(np.zeros((3, 7), dtype=np.float32))
sim.close()
print('ASAN-WORKLOAD-OK')
