#!/usr/bin/env python3
"""GPU box: BedBathingSawyer-v1 fused step throughput, device-timed; not the headline metric."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from assistive_gym_b200 import capi
from assistive_gym_b200.bed_bathing_batch import BedBathingBatch
from assistive_gym_b200.sim import BatchSim
n = int(os.environ.get('AG_N', '4096')); K = int(os.environ.get('AG_STEPS', '20'))
bb = BedBathingBatch(); sim = BatchSim(bb.scene, capi.default_config(), n)
t0 = time.time(); s = bb.reset(sim, np.random.default_rng(0)); bb.start_fused(sim, s)
print('reset %.1f s, unresolved start poses %d of %d, base draws %d' % (time.time() - t0, bb.unresolved, n, bb.base_draws))
stream = torch.cuda.ExternalStream(sim.stream_ptr())
dev = torch.device('cuda')
act = torch.rand((K + 5, n, 7), device=dev) * 2 - 1
obs = torch.zeros((n, 24), device=dev); rew = torch.zeros(n, device=dev); done = torch.zeros(n, device=dev); info = torch.zeros((n, 4), device=dev)
torch.cuda.synchronize()
for i in range(5): sim.bathing_step_dev(act[i].data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(), info.data_ptr())
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream): a.record(stream)
for i in range(K): sim.bathing_step_dev(act[5 + i].data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(), info.data_ptr())
with torch.cuda.stream(stream): b.record(stream)
torch.cuda.synchronize()
ms = a.elapsed_time(b) / K
cnt, it = sim.solver_stats()
print('BedBathingSawyer-v1 batch %d: %.2f ms/step, %.0f env-steps/s; contacts/env mean %.1f, PGS iters mean %.1f; wiped so far %d, cloth force max %.1f N'
      % (n, ms, n / ms * 1e3, cnt.mean(), it.mean(), int(info[:, 3].sum().item()), float(info[:, 2].max().item())))
