#!/usr/bin/env python3
"""Diagnostic (GPU box): per-env PGS cycles vs contact count / iterations after a few steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from assistive_gym_b200 import capi
from assistive_gym_b200.feeding_batch import FeedingBatch
from assistive_gym_b200.sim import BatchSim
n = int(os.environ.get("AG_N", "4096"))
fb = FeedingBatch(); sim = BatchSim(fb.scene, capi.default_config(), n)
rng = np.random.default_rng(0)
s = fb.reset(sim, rng, settle_steps=25)
fb.start_fused(sim, s)
for i in range(8):
    sim.feeding_step_host(rng.uniform(-1, 1, size=(n, 7)).astype(np.float32))
cyc = sim.pgs_cycles().astype(np.float64); cnt, it = sim.solver_stats()
print('cycles: mean %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f' % (cyc.mean(), *np.percentile(cyc, [50, 90, 99]), cyc.max()))
o = np.argsort(-cyc)[:12]
for e in o:
    c, nc = sim.get_contacts(fb.robot, max_pts=1)
    print('env %4d cycles %9.0f contacts %3d iters %2d robot-contacts %d  cycles/(iter*contact) %.0f' % (e, cyc[e], cnt[e], it[e], nc[e], cyc[e] / max(1, it[e] * cnt[e])))
work = it * (3 * cnt + 46)
print('corr(cycles, iters*rows) = %.3f ; cycles per row-update: median %.0f' % (np.corrcoef(cyc, work)[0, 1], np.median(cyc / np.maximum(work, 1))))
