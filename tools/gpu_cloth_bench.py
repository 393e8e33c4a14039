"""k_cloth on the reference's gown (3 966 nodes, numSubSteps = 8) over the synthetic obstacle scene of tests/cloth_cases.py:
per-launch device time and achieved algorithmic bandwidth.  `python tools/gpu_cloth_bench.py [N] [steps]`"""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from assistive_gym_b200 import capi
from assistive_gym_b200.cloth import ClothModel
from assistive_gym_b200.sim import BatchSim
from tests import cloth_cases as cc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
model = ClothModel.load()
scene, links, static, arm_joint = cc.obstacle_scene()
cfg = capi.default_config(num_substeps=8)
sim = BatchSim(scene, cfg, n)
anchors = [2086, 2087, 2088, 2041]
rng = np.random.default_rng(0)
x0 = np.repeat(model.rest[None], n, axis=0).astype(np.float32)
x0 = x0 - x0.mean(axis=1, keepdims=True) + np.array([0.2, 0.15, 0.38], dtype=np.float32)
x0[:, :, :2] += rng.uniform(-0.03, 0.03, size=(n, 1, 2)).astype(np.float32)
sim.cloth_init(model, links, static, anchors, model.rest[anchors] - model.rest[anchors[0]], max_contacts=1024)
sim.cloth_set_state(x0, np.zeros_like(x0))
sim.cloth_set_anchor(x0[:, anchors[0]].copy())
sim.set_joint_state([arm_joint], q=np.full((n, 1), -0.8), qd=np.full((n, 1), 2.0))
sim.forward_kinematics()
sim.step(3)
sim.profile_enable(True)
sim.step(steps)
prof = sim.profile_get()
sim.profile_enable(False)
cnt = sim.cloth_get_contacts(16)[0]
ms = prof['k_cloth'][0] / prof['k_cloth'][1]
alg = n * 8 * model.n_nodes * 6 * 4 * 2            # x, v read + written per substep (190 KB per env-substep)
print(json.dumps({'n_envs': n, 'k_cloth_ms_per_launch': ms, 'launches': prof['k_cloth'][1],
                  'algorithmic_GBps': alg / ms / 1e6, 'contacts_mean': float(cnt.mean()), 'contacts_max': int(cnt.max()),
                  'overflow_envs': sim.overflow_count(),
                  'all_kernels_ms': {k: v[0] / steps for k, v in prof.items()}}))
