#!/usr/bin/env python3
"""Generates tests/golden/feeding_10substeps.npz from the CPU oracle (NOT from PyBullet: the
reference's physics cannot run here — parity unpinned).  Committed together with its output."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from assistive_gym_b200 import capi  # noqa: E402
from assistive_gym_b200.feeding_batch import FeedingBatch  # noqa: E402
from oracle.oracle_py import OracleSim  # noqa: E402

fb = FeedingBatch()
n = 4
cfg = capi.default_config(residual_threshold=0.0)
cpu = OracleSim(fb.scene, cfg, n)
s = fb.reset(cpu, np.random.default_rng(42), settle_steps=25)
state0 = cpu.state_get()
q = cpu.get_joint_states(fb.arm_links)[0]
targets = q + np.random.default_rng(43).uniform(-0.2, 0.2, size=q.shape)
cpu.set_motor_targets(fb.arm_links, targets)
cpu.step(10)
out = dict(state0=state0, state10=cpu.state_get(), targets=targets)
out.update({'s_' + k: v for k, v in s.items()})
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'feeding_10substeps.npz'), **out)
print('wrote fixture', {k: v.shape for k, v in out.items()})
