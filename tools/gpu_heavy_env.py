#!/usr/bin/env python3
"""Diagnostic (GPU box): what do the slowest PGS envs look like?  Contact counts per body pair."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from assistive_gym_b200 import capi
from assistive_gym_b200.feeding_batch import FeedingBatch
from assistive_gym_b200.sim import BatchSim
n = int(os.environ.get("AG_N", "4096")); steps = int(os.environ.get("AG_STEPS", "30"))
fb = FeedingBatch(); sim = BatchSim(fb.scene, capi.default_config(), n)
from assistive_gym_b200.sharding import sample_block
lo = int(os.environ.get("AG_LO", "-1"))
if lo >= 0:      # the bench's shard [lo, lo + n) (rank = lo / n)
    rng = np.random.default_rng(1001 + lo)
    s = fb.reset(sim, rng, settle_steps=25, sample=sample_block(fb, lo, lo + n))
    print('shard', lo, 'ik_colliding', fb.ik_colliding, 'ik_resamples', fb.ik_resamples, 'ik_err max', float(fb.ik_err.max()))
else:
    rng = np.random.default_rng(0)
    s = fb.reset(sim, rng, settle_steps=25)
fb.start_fused(sim, s)
import torch
acts = (torch.rand((steps, n, 7), generator=torch.Generator().manual_seed(lo // n if lo >= 0 else 0)) * 2 - 1).numpy()
for i in range(steps):
    sim.feeding_step_host(acts[i])
cyc = sim.pgs_cycles().astype(np.float64); cnt, it = sim.solver_stats()
print('cycles: mean %.0f p50 %.0f p90 %.0f p99 %.0f p99.9 %.0f max %.0f' % (cyc.mean(), *np.percentile(cyc, [50, 90, 99, 99.9]), cyc.max()))
print('iters==50: %.3f of envs; contacts mean %.1f max %d' % ((it >= 50).mean(), cnt.mean(), cnt.max()))
sc = fb.scene
names = {fb.plane: 'plane', fb.robot: 'robot', fb.humans['male']: 'human_m', fb.humans['female']: 'human_f', fb.wheelchair: 'wheelchair',
         fb.table: 'table', fb.tool: 'tool', fb.bowl: 'bowl'}
for f in fb.foods: names[f] = 'food'
order = np.argsort(-cyc)[:6]
allc = {}
for b in range(sc.n_bodies):
    allc[b] = sim.get_contacts(b, max_pts=128)
for e in order:
    pairs = {}
    for b in range(sc.n_bodies):
        c, k = allc[b]
        for j in range(k[e]):
            ba, bb = int(sc['link_body'][c[e, j]['link_a']]), int(sc['link_body'][c[e, j]['link_b']])
            if ba != b: continue
            key = tuple(sorted((names.get(ba, str(ba)), names.get(bb, str(bb)))))
            pairs[key] = pairs.get(key, 0) + 1
    print('env %d cycles %.2fM iters %d contacts %d impairment %d :' % (e, cyc[e] / 1e6, it[e], cnt[e], s['impairment'][e]), {k: v for k, v in sorted(pairs.items(), key=lambda x: -x[1])})
