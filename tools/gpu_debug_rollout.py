#!/usr/bin/env python3
"""Debug aid (GPU box): per-env-step error of the CUDA build vs the oracle on the strict rollout."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from assistive_gym_b200 import capi
from assistive_gym_b200.feeding_batch import FeedingBatch
from assistive_gym_b200.sim import BatchSim
from tests import parity_cases as pc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fb = FeedingBatch()
mk = lambda scene, cfg, n: BatchSim(scene, cfg, n, device=0)
cfg = capi.default_config(residual_threshold=0.0)
cpu, dev, _ = pc.synced_pair(fb, mk, n, 0, cfg)
for f in fb.foods:
    cpu.set_body_active(f, 0); dev.set_body_active(f, 0)
rng = np.random.default_rng(100)
for k in range(steps):
    act = rng.uniform(-1, 1, size=(n, 7))
    tgt = pc.take_step_targets(cpu.get_joint_states(fb.arm_links)[0], act, fb.arm_lower, fb.arm_upper)
    cpu.set_motor_targets(fb.arm_links, tgt); dev.set_motor_targets(fb.arm_links, tgt)
    for s in range(5):
        st = cpu.state_get()
        cpu.step(1); dev.step(1)
        d = np.abs(cpu.get_joint_states(fb.arm_links + fb.gripper_links)[0] - dev.get_joint_states(fb.arm_links + fb.gripper_links)[0])
        if d.max() > 1e-5:
            e = int(d.max(axis=1).argmax())
            print('step', k, s, 'env', e, 'dq', np.array2string(d[e], precision=1), 'contacts cpu/dev', cpu.get_contacts(fb.robot, max_pts=1)[1][e], dev.get_contacts(fb.robot, max_pts=1)[1][e],
                  'ncon', cpu.num_contacts()[0][e], dev.get_contacts(fb.tool, max_pts=1)[1][e])
            ca, na = cpu.get_contacts(fb.robot, max_pts=8); cc, nc = dev.get_contacts(fb.robot, max_pts=8)
            print('  cpu', [(int(c['link_a']), int(c['link_b']), float(c['distance']), float(c['normal_force'])) for c in ca[e, :na[e]]])
            print('  dev', [(int(c['link_a']), int(c['link_b']), float(c['distance']), float(c['normal_force'])) for c in cc[e, :nc[e]]])
            qa, qda, ta = cpu.get_joint_states(fb.arm_links + fb.gripper_links); qc, qdc, tc = dev.get_joint_states(fb.arm_links + fb.gripper_links)
            print('  qd cpu', np.round(qda[e], 4)); print('  qd dev', np.round(qdc[e], 4)); print('  tau cpu', np.round(ta[e], 4)); print('  tau dev', np.round(tc[e], 4))
            sys.exit(0)
    if k % 10 == 0:
        print('step', k, 'max dq', d.max())
print('no divergence above 1e-5')
