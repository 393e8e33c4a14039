#!/usr/bin/env python3
"""Asset compiler: reference URDF / mesh assets -> compact `.agmodel.json` model files.

The reference ships 216 MB of URDF/OBJ/DAE/STL assets that do not exist on
the GPU box.  This script reads the ones the hot path needs from a reference checkout and writes
small derived model files (kinematic tree + inertials + convex-hull collider vertices) under
`assistive_gym_b200/assets/`.  The outputs are committed; this script is the committed generator.

What it restates of PyBullet's URDF import (not verifiable here):
  * link index == joint index == DFS pre-order over the URDF tree, children in file order
    (checked against reference agents/jaco.py:8-18, sawyer.py:8-17, pr2.py:8-18 in tests),
  * mesh colliders become convex hulls: one hull per `o` group for .obj, one hull per file for
    .dae/.stl; COLLADA <up_axis>/<unit> honoured,
  * hulls with more than MAX_HULL_VERTS vertices are simplified (greedy farthest-point) and the
    geometric error is recorded in the output (`hull_simplify_err_m`).

Usage: python tools/compile_assets.py [--ref /root/reference] [--out assistive_gym_b200/assets]
"""
import argparse
import json
import os
import sys
import struct
import xml.etree.ElementTree as ET

import numpy as np
from scipy.spatial import ConvexHull

MAX_HULL_VERTS = 64


# ----------------------------------------------------------------------------- math helpers
def rpy_to_mat(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def fl(s, n=3, default=0.0):
    if s is None:
        return [default] * n
    return [float(x) for x in s.split()]


# ----------------------------------------------------------------------------- mesh readers
def load_obj_groups(path):
    groups, cur = [], None
    for line in open(path, errors='ignore'):
        if line.startswith('o '):
            cur = []
            groups.append(cur)
        elif line.startswith('v '):
            if cur is None:
                cur = []
                groups.append(cur)
            cur.append([float(x) for x in line.split()[1:4]])
    return [np.array(g, dtype=np.float64) for g in groups if len(g) >= 4]


def load_stl(path):
    data = open(path, 'rb').read()
    ntri = struct.unpack('<I', data[80:84])[0]
    if 84 + 50 * ntri == len(data):
        arr = np.frombuffer(data, dtype=np.dtype([('n', '<f4', 3), ('v', '<f4', (3, 3)), ('a', '<u2')]),
                            count=ntri, offset=84)
        return [arr['v'].reshape(-1, 3).astype(np.float64)]
    verts = []
    for line in data.decode(errors='ignore').splitlines():
        t = line.split()
        if len(t) == 4 and t[0] == 'vertex':
            verts.append([float(x) for x in t[1:]])
    return [np.array(verts)]


def load_dae(path):
    ns = {'c': 'http://www.collada.org/2005/11/COLLADASchema'}
    root = ET.parse(path).getroot()
    up = root.find('c:asset/c:up_axis', ns)
    unit = root.find('c:asset/c:unit', ns)
    scale = float(unit.get('meter', '1')) if unit is not None else 1.0
    pts = []
    for geom in root.findall('c:library_geometries/c:geometry', ns):
        mesh = geom.find('c:mesh', ns)
        if mesh is None:
            continue
        vin = mesh.find('c:vertices/c:input[@semantic="POSITION"]', ns)
        src = vin.get('source')[1:]
        for s in mesh.findall('c:source', ns):
            if s.get('id') == src:
                a = np.array([float(x) for x in s.find('c:float_array', ns).text.split()])
                pts.append(a.reshape(-1, 3))
    v = np.concatenate(pts) * scale
    upv = up.text.strip() if up is not None else 'Y_UP'
    if upv == 'Y_UP':      # rotate +90deg about X: (x, y, z) -> (x, -z, y)
        v = np.stack([v[:, 0], -v[:, 2], v[:, 1]], axis=1)
    elif upv == 'X_UP':    # rotate -90deg about Y
        v = np.stack([-v[:, 2], v[:, 1], v[:, 0]], axis=1)
    return [v]


def load_mesh_groups(path):
    ext = os.path.splitext(path)[1].lower()
    if ext == '.obj':
        return load_obj_groups(path)
    if ext == '.dae':
        return load_dae(path)
    if ext == '.stl':
        return load_stl(path)
    raise ValueError('unsupported mesh ' + path)


# ----------------------------------------------------------------------------- hull tools
def hull_planes(verts):
    """Unique outward face planes (n, d) with n.x <= d inside; merged coplanar facets."""
    h = ConvexHull(verts)
    eq = h.equations  # n.x + off <= 0 inside
    planes = []
    for e in eq:
        n, d = e[:3], -e[3]
        dup = False
        for (pn, pd) in planes:
            if np.dot(pn, n) > 1.0 - 1e-6 and abs(pd - d) < 1e-7:
                dup = True
                break
        if not dup:
            planes.append((n, d))
    return planes


def simplify_hull(points, max_verts=MAX_HULL_VERTS):
    """Return (hull vertices <= max_verts, max distance of dropped hull vertices outside the kept hull)."""
    h = ConvexHull(points)
    hv = points[h.vertices]
    if len(hv) <= max_verts:
        return hv, 0.0
    # greedy: start from axis extremes, add the vertex farthest outside the current hull
    keep = set()
    for ax in range(3):
        keep.add(int(np.argmin(hv[:, ax])))
        keep.add(int(np.argmax(hv[:, ax])))
    keep = list(keep)
    while len(keep) < 4:
        keep.append(int(np.setdiff1d(np.arange(len(hv)), keep)[0]))
    err = 0.0
    while True:
        cur = ConvexHull(hv[keep], qhull_options='QJ')
        dist = (hv @ cur.equations[:, :3].T + cur.equations[:, 3]).max(axis=1)
        dist[keep] = -1.0
        j = int(np.argmax(dist))
        err = max(float(dist[j]), 0.0)
        if len(keep) >= max_verts or err <= 1e-9:
            break
        keep.append(j)
    sub = hv[keep]
    sub = sub[ConvexHull(sub, qhull_options='QJ').vertices]
    return sub, err


# ----------------------------------------------------------------------------- URDF -> model
def parse_collider(col, urdf_dir, stats):
    o = col.find('origin')
    xyz = fl(o.get('xyz') if o is not None else None)
    rpy = fl(o.get('rpy') if o is not None else None)
    g = col.find('geometry')[0]
    out = {'origin_xyz': xyz, 'origin_rpy': rpy}
    if g.tag == 'box':
        out.update(type='box', size=fl(g.get('size')))
    elif g.tag == 'sphere':
        out.update(type='sphere', radius=float(g.get('radius')))
    elif g.tag in ('cylinder', 'capsule'):
        out.update(type=g.tag, radius=float(g.get('radius')), length=float(g.get('length')))
    elif g.tag == 'mesh':
        scale = fl(g.get('scale'), default=1.0) if g.get('scale') else [1.0, 1.0, 1.0]
        fn = g.get('filename')
        hulls = compile_mesh_hulls(os.path.join(urdf_dir, fn), scale, stats)
        out.update(type='mesh', file=os.path.basename(fn), scale=scale, hulls=hulls)
    else:
        raise ValueError(g.tag)
    return out


def compile_mesh_hulls(path, scale, stats):
    hulls = []
    for grp in load_mesh_groups(path):
        pts = np.unique(np.round(grp * np.array(scale), 9), axis=0)
        hv, err = simplify_hull(pts)
        stats['max_err'] = max(stats.get('max_err', 0.0), err)
        stats['hulls'] = stats.get('hulls', 0) + 1
        stats['verts'] = stats.get('verts', 0) + len(hv)
        hulls.append(np.round(hv, 7).tolist())
    return hulls


def compile_urdf(urdf_path):
    root = ET.parse(urdf_path).getroot()
    urdf_dir = os.path.dirname(urdf_path)
    links = {l.get('name'): l for l in root.findall('link')}
    joints = root.findall('joint')
    children = {}
    child_names = set()
    for j in joints:
        children.setdefault(j.find('parent').get('link'), []).append(j)
        child_names.add(j.find('child').get('link'))
    roots = [n for n in links if n not in child_names]
    assert len(roots) == 1, roots
    stats = {}
    out_links = []

    def emit(name, parent_idx, joint):
        l = links[name]
        rec = {'name': name, 'parent': parent_idx}
        if joint is not None:
            o = joint.find('origin')
            ax = joint.find('axis')
            lim = joint.find('limit')
            dyn = joint.find('dynamics')
            rec['joint'] = {
                'name': joint.get('name'), 'type': joint.get('type'),
                'origin_xyz': fl(o.get('xyz') if o is not None else None),
                'origin_rpy': fl(o.get('rpy') if o is not None else None),
                'axis': fl(ax.get('xyz')) if ax is not None else [1.0, 0.0, 0.0],
                'lower': float(lim.get('lower', 0)) if lim is not None else 0.0,
                'upper': float(lim.get('upper', 0)) if lim is not None else 0.0,
                'effort': float(lim.get('effort', 0)) if lim is not None else 0.0,
                'velocity': float(lim.get('velocity', 0)) if lim is not None else 0.0,
                'damping': float(dyn.get('damping', 0)) if dyn is not None else 0.0,
            }
        iner = l.find('inertial')
        if iner is not None:
            o = iner.find('origin')
            I = iner.find('inertia')
            rec['inertial'] = {
                'mass': float(iner.find('mass').get('value')),
                'com_xyz': fl(o.get('xyz') if o is not None else None),
                'com_rpy': fl(o.get('rpy') if o is not None else None),
                'inertia': [float(I.get(k, 0)) for k in ('ixx', 'iyy', 'izz', 'ixy', 'ixz', 'iyz')] if I is not None else [0] * 6,
            }
        else:
            rec['inertial'] = {'mass': 0.0, 'com_xyz': [0, 0, 0], 'com_rpy': [0, 0, 0], 'inertia': [0] * 6}
        c = l.find('contact')
        rec['contact'] = {}
        if c is not None:
            for tag in ('lateral_friction', 'rolling_friction', 'spinning_friction'):
                e = c.find(tag)
                if e is not None:
                    rec['contact'][tag] = float(e.get('value'))
        rec['colliders'] = [parse_collider(col, urdf_dir, stats) for col in l.findall('collision')]
        idx = len(out_links)
        out_links.append(rec)
        for j in children.get(name, []):
            emit(j.find('child').get('link'), idx, j)

    emit(roots[0], -1, None)
    return {'name': root.get('name'), 'source': os.path.basename(urdf_path), 'links': out_links,
            'hull_simplify_err_m': stats.get('max_err', 0.0), 'n_hulls': stats.get('hulls', 0),
            'n_hull_verts': stats.get('verts', 0), 'max_hull_verts': MAX_HULL_VERTS}


def compile_cloth(path):
    """Triangle mesh of a cloth (reference dressing.py:146 `p.loadCloth(...hospitalgown_reduced.obj...)`).

    Node numbering restates the old tinyobjloader bundled with Bullet (recalled): a node is created for every distinct
    `v/vt/vn` corner triple, numbered in order of FIRST APPEARANCE in the face list -- not in `v` line order.  This is
    pinned by the reference's own constants (tests/test_cloth_model.py): with this numbering the anchor nodes
    [2086, 2087, 2088, 2041] (dressing.py:146) sit within 2.4 cm of `cloth_orig_pos` (dressing.py:140) and the six sleeve
    nodes (dressing.py:149-150) form a 15 cm ring; with `v` line order they are scattered over the gown.
    """
    V, F = [], []
    for line in open(path, errors='ignore'):
        if line.startswith('v '):
            V.append([float(x) for x in line.split()[1:4]])
        elif line.startswith('f '):
            F.append(line.split()[1:4])
    V = np.array(V, dtype=np.float64)
    ids, order, faces = {}, [], []
    for f in F:
        t = []
        for c in f:
            if c not in ids:
                ids[c] = len(order)
                order.append(int(c.split('/')[0]) - 1)
            t.append(ids[c])
        faces.append(t)
    return V[np.array(order)], np.array(faces, dtype=np.int32)


def compile_mesh_asset(path, name):
    stats = {}
    hulls = compile_mesh_hulls(path, [1.0, 1.0, 1.0], stats)
    return {'name': name, 'source': os.path.basename(path), 'hulls': hulls,
            'hull_simplify_err_m': stats.get('max_err', 0.0), 'n_hulls': len(hulls),
            'n_hull_verts': stats.get('verts', 0)}


URDFS = {
    'plane': 'plane/plane.urdf',
    'jaco': 'jaco/j2s7s300_gym.urdf',
    'wheelchair_jaco': 'wheelchair/wheelchair_jaco.urdf',
    'wheelchair': 'wheelchair/wheelchair.urdf',
    'table_tall': 'table/table_tall.urdf',
    'bowl': 'dinnerware/bowl.urdf',
    'sawyer': 'sawyer/sawyer.urdf',
    'bed': 'bed/bed.urdf',
    'wiper': 'bed_bathing/wiper.urdf',
    'pr2': 'PR2/pr2_no_torso_lift_tall.urdf',
    'tool_scratch': 'scratcher/tool_scratch.urdf',
}
CLOTHS = {
    'hospitalgown_reduced': 'clothing/hospitalgown_reduced.obj',
}
MESHES = {
    'spoon_vhacd': 'dinnerware/spoon_vhacd.obj',
    'cup_vhacd': 'dinnerware/plastic_coffee_cup_vhacd.obj',
    'head_male_vhacd': 'head_female_male/BaseHeadMeshes_v5_male_cropped_reduced_compressed_vhacd.obj',
    'head_female_vhacd': 'head_female_male/BaseHeadMeshes_v5_female_cropped_reduced_compressed_vhacd.obj',
}


def compile_keras_mlp(path):
    """Dense layers (kernel [in][out], bias [out], activation) of a Keras Sequential model saved as HDF5: the reference's
    realistic joint-limit classifier (envs/env.py:39, agents/human.py:134-152), 4 -> 64 -> 64 -> 64 -> 1."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from minih5 import MiniH5
    h = MiniH5(path)
    w = {k: v for k, v in h.walk().items() if k.startswith('/model_weights/')}
    raw = open(path, 'rb').read()
    i = raw.find(b'{"class_name": "Sequential"')
    cfg = json.loads(raw[i:raw.find(b'\0', i)].decode())
    layers = cfg['config'] if isinstance(cfg['config'], list) else cfg['config']['layers']
    out = {}
    for n, l in enumerate(layers):
        assert l['class_name'] == 'Dense' and l['config']['use_bias']
        name = l['config']['name']
        out['W%d' % n] = w['/model_weights/%s/%s/kernel:0' % (name, name)].astype(np.float32)
        out['b%d' % n] = w['/model_weights/%s/%s/bias:0' % (name, name)].astype(np.float32)
        out['act%d' % n] = np.array(l['config']['activation'])
    out['n_layers'] = np.array(len(layers))
    return out


MLPS = {'realistic_arm_limits_model': 'realistic_arm_limits_model.h5'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference')
    ap.add_argument('--out', default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                  'assistive_gym_b200', 'assets'))
    ap.add_argument('--only', default=None)
    args = ap.parse_args()
    adir = os.path.join(args.ref, 'assistive_gym', 'envs', 'assets')
    os.makedirs(args.out, exist_ok=True)
    for name, rel in URDFS.items():
        if args.only and name != args.only:
            continue
        p = os.path.join(adir, rel)
        if not os.path.exists(p):
            print('skip (missing)', rel)
            continue
        try:
            m = compile_urdf(p)
        except Exception as e:  # noqa
            print('FAILED', name, repr(e))
            continue
        json.dump(m, open(os.path.join(args.out, name + '.agmodel.json'), 'w'), separators=(',', ':'))
        print('%-16s links=%d hulls=%d verts=%d simplify_err=%.2e m' % (name, len(m['links']), m['n_hulls'], m['n_hull_verts'], m['hull_simplify_err_m']))
    for name, rel in MESHES.items():
        if args.only and name != args.only:
            continue
        m = compile_mesh_asset(os.path.join(adir, rel), name)
        json.dump(m, open(os.path.join(args.out, name + '.agmesh.json'), 'w'), separators=(',', ':'))
        print('%-16s hulls=%d verts=%d simplify_err=%.2e m' % (name, m['n_hulls'], m['n_hull_verts'], m['hull_simplify_err_m']))


    for name, rel in CLOTHS.items():
        if args.only and name != args.only:
            continue
        v, f = compile_cloth(os.path.join(adir, rel))
        np.savez_compressed(os.path.join(args.out, name + '.agcloth.npz'), verts=v, faces=f)
        print('%-16s cloth nodes=%d faces=%d' % (name, len(v), len(f)))
    for name, rel in MLPS.items():
        if args.only and name != args.only:
            continue
        m = compile_keras_mlp(os.path.join(adir, rel))
        np.savez_compressed(os.path.join(args.out, name + '.agmlp.npz'), **m)
        print('%-16s dense layers: %s' % (name, ' -> '.join([str(m['W0'].shape[0])] + [str(m['W%d' % k].shape[1]) for k in range(int(m['n_layers']))])))


if __name__ == '__main__':
    main()
