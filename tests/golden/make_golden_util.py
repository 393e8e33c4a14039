"""Golden vectors of the reference's geometric helpers (assistive_gym/envs/util.py), recorded by running the reference's own code in
this container (its `import pybullet` is satisfied by an empty stub: the functions used here are plain numpy):
  capsule_points (util.py:80-113, the wiping targets of BedBathing), points_in_cylinder (:53-56, Drinking),
  line_intersects_triangle (:125-132) and sleeve_on_arm_reward (:134-202, Dressing) on seeded random inputs.
Written to tests/golden/util_vectors.npz; tests/test_reference_util_vectors.py compares the repo's restatements against it.

usage: python tests/golden/make_golden_util.py [/root/reference]"""
import importlib.util
import os
import sys
import types

import numpy as np


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    sys.modules['pybullet'] = types.ModuleType('pybullet')
    spec = importlib.util.spec_from_file_location('ref_util', os.path.join(ref, 'assistive_gym', 'envs', 'util.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    del sys.modules['pybullet']
    u = mod.Util(0, np.random.RandomState(0))
    rng = np.random.default_rng(0)
    out = {}
    # capsule_points: the limb capsules of BedBathing (bed_bathing.py:176-187 uses p1 = 0, p2 = (0, 0, -length), several radii) + skew axes
    cp_in, cp_out, cp_n = [], [], []
    cases = [((0, 0, 0), (0, 0, -0.279), 0.043), ((0, 0, 0), (0, 0, -0.257), 0.033), ((0, 0, 0), (0, 0, -0.264), 0.0355), ((0, 0, 0), (0, 0, -0.234), 0.027)]
    for _ in range(6):
        cases.append((rng.normal(size=3), rng.normal(size=3), float(rng.uniform(0.02, 0.08))))
    for p1, p2, r in cases:
        for d in (0.03, 0.05):
            pts = np.array(u.capsule_points(p1=np.array(p1, dtype=float), p2=np.array(p2, dtype=float), radius=r, distance_between_points=d)).reshape(-1, 3)
            cp_in.append(list(p1) + list(p2) + [r, d]); cp_n.append(len(pts)); cp_out.append(pts)
    out['capsule_in'], out['capsule_n'], out['capsule_pts'] = np.array(cp_in, dtype=np.float64), np.array(cp_n), np.concatenate(cp_out)
    # points_in_cylinder
    pc_in = rng.normal(size=(200, 10)); pc_in[:, 6] = np.abs(pc_in[:, 6]) * 0.5 + 0.05
    pc_in[:100, 7:10] = pc_in[:100, 0:3] + (pc_in[:100, 3:6] - pc_in[:100, 0:3]) * rng.uniform(-0.2, 1.2, size=(100, 1)) + rng.normal(size=(100, 3)) * 0.1
    out['cyl_in'] = pc_in
    out['cyl_out'] = np.array([bool(u.points_in_cylinder(r[0:3], r[3:6], r[6], r[7:10])) for r in pc_in])
    # line_intersects_triangle
    lt = rng.normal(size=(400, 15))
    out['tri_in'] = lt
    out['tri_out'] = np.array([bool(u.line_intersects_triangle(r[0:3], r[3:6], r[6:9], r[9:12], r[12:15])) for r in lt])
    # sleeve_on_arm_reward: an arm (shoulder - elbow - wrist) and two triangles on a ring of random radius at a random place along it
    sl_in, sl_out = [], []
    for k in range(400):
        shoulder = rng.normal(size=3) * 0.1 + np.array([0, 0, 1.0])
        elbow = shoulder + rng.normal(size=3) * 0.05 + np.array([0, -0.28, 0]) * rng.uniform(0.8, 1.2)
        wrist = elbow + rng.normal(size=3) * 0.05 + np.array([0, -0.26, 0.0]) * rng.uniform(0.8, 1.2)
        seg = rng.integers(0, 3)
        a, b_ = (elbow, wrist) if seg == 0 else (shoulder, elbow) if seg == 1 else (wrist, wrist + (wrist - elbow))
        c = a + (b_ - a) * rng.uniform(-0.1, 1.1) + rng.normal(size=3) * rng.choice([0.005, 0.05, 0.2])
        ax = (b_ - a) / np.linalg.norm(b_ - a)
        e1 = np.cross(ax, [0.3, 0.5, 0.8]); e1 /= np.linalg.norm(e1); e2 = np.cross(ax, e1)
        rad = rng.uniform(0.05, 0.12)
        ang = np.sort(rng.uniform(0, 2 * np.pi, size=6))
        ring = np.array([c + rad * (np.cos(t) * e1 + np.sin(t) * e2) + ax * rng.normal() * 0.01 for t in ang])
        t1, t2 = ring[[0, 2, 4]], ring[[1, 3, 5]]
        radii = rng.uniform(0.02, 0.05, size=3)
        res = u.sleeve_on_arm_reward(t1, t2, shoulder, elbow, wrist, radii[0], radii[1], radii[2])
        sl_in.append(np.concatenate([t1.ravel(), t2.ravel(), shoulder, elbow, wrist, radii]))
        sl_out.append([float(v) for v in res])
    out['sleeve_in'], out['sleeve_out'] = np.array(sl_in), np.array(sl_out)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'util_vectors.npz')
    np.savez_compressed(path, **out)
    so = out['sleeve_out']
    print('capsule cases', len(cp_n), 'points', int(np.sum(cp_n)), '| cylinder inside', int(out['cyl_out'].sum()), '/ 200 | triangle hits', int(out['tri_out'].sum()),
          '/ 400 | sleeve: forearm', int(so[:, 0].sum()), 'upperarm', int(so[:, 1].sum()), 'of 400')


if __name__ == '__main__':
    main()
