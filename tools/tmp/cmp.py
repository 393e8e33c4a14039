import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import numpy as np
from assistive_gym_b200 import capi
from assistive_gym_b200.feeding_batch import FeedingBatch
from assistive_gym_b200.sim import BatchSim
emu = capi.load_library(os.path.join(ROOT, 'tests', 'kernel_harness', 'libagphys_emu.so'))
gpu = capi.load_library()
fb = FeedingBatch(); n = 8
cfg = capi.default_config(residual_threshold=0.0)
a = BatchSim(fb.scene, cfg, n, _lib=emu); b = BatchSim(fb.scene, cfg, n, _lib=gpu)
s = fb.reset(a, np.random.default_rng(2), settle_steps=5)
fb.reset(b, np.random.default_rng(2), settle_steps=0, sample=s)
b.state_set(a.state_get())
q = a.get_joint_states(fb.arm_links)[0]
for sim in (a, b):
    sim.set_motor_targets(fb.arm_links, q)
mode = sys.argv[1] if len(sys.argv) > 1 else 'tool'
if mode == 'tool':
    head = {1: fb.gl(fb.humans['male'], 23), 0: fb.gl(fb.humans['female'], 23)}
    hl = np.array([head[int(m)] for m in s['male']])
    hp = np.stack([a.get_link_states([int(h)])['pos'][e, 0] for e, h in enumerate(hl)])
    for sim in (a, b):
        sim.set_body_active(fb.robot, 0)
        for f in fb.foods: sim.set_body_active(f, 0)
        sim.set_base_pose(fb.tool, hp + np.array([0.0, -0.25, 0.0]), np.array([0.7071068, 0, 0, 0.7071068]))
        sim.set_base_velocity(fb.tool, np.tile([0.0, 0.6, 0.0], (n, 1)), np.zeros((n, 3)))
        sim.forward_kinematics()
else:
    for sim in (a, b):
        for f in fb.foods: sim.set_body_active(f, 0)
for i in range(12):
    a.step(1); b.step(1)
    sa, sb = a.state_get(), b.state_get()
    d = np.abs(sa - sb).max(axis=1)
    t, f = b.pgs_trips(); ca, ia = a.solver_stats(); cb, ib = b.solver_stats()
    print('step', i, 'max state diff per env', np.array2string(d, precision=2), 'floats', f, 'contacts', ca, cb, 'iters', ia, ib)
    b.state_set(sa)
