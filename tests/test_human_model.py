"""`assistive_gym_b200/human_model.py` (the capsule person, restated from reading reference envs/human_creation.py) against what the
reference's own code asks the physics engine to create: tests/golden/human_<gender>.json, recorded by running
`HumanCreation.create_human` against a stub pybullet (tests/golden/make_golden_human.py).  Link numbering: PyBullet numbers the
links of `createMultiBody` depth-first in creation order; the reference's joint legend (human_creation.py:5-47)
is that numbering, and so is the scene's."""
import json
import os

import numpy as np
import pytest

from assistive_gym_b200.human_model import create_human
from assistive_gym_b200.scene import SceneBuilder

HERE = os.path.dirname(os.path.abspath(__file__))
COL_SPHERE, COL_CAPSULE = 0, 1


def _qrot(q, v):
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return R @ np.asarray(v, dtype=np.float64)


def _dfs_order(parents):
    kids = {i: [] for i in range(len(parents) + 1)}
    for i, p in enumerate(parents):
        kids[int(p)].append(i + 1)
    order = []

    def walk(k):
        for c in kids[k]:
            order.append(c)
            walk(c)
    walk(0)
    return order


@pytest.mark.parametrize('gender', ['male', 'female'])
def test_person_matches_what_the_reference_creates(gender):
    g = json.load(open(os.path.join(HERE, 'golden', 'human_%s.json' % gender)))
    mb, shapes = g['multibody'], g['shapes']
    b = SceneBuilder()
    hb, info = create_human(b, gender=gender, static=True)
    sc = b.finalize()
    l0, n = int(sc['body_link0'][hb]), int(sc['body_nlinks'][hb])
    assert n == g['n_links'] + 1 == 43
    order = _dfs_order(mb['linkParentIndices'])                     # new index -> creation index (1-based)
    new_of = {old: new for new, old in enumerate(order)}
    col_link = np.asarray(sc['col_link'])

    def check_shape(link, sh):
        cols = np.where(col_link == link)[0]
        if sh is None:
            assert len(cols) == 0
            return
        if sh['kind'] == 'mesh':                                     # the head: a convex decomposition of the reference's mesh file
            assert len(cols) >= 1
            return
        assert len(cols) == 1
        c = int(cols[0])
        v0, nv = int(sc['col_v0'][c]), int(sc['col_nv'][c])
        verts = np.asarray(sc['verts'])[v0:v0 + nv]
        assert abs(float(sc['col_radius'][c]) - sh['radius']) < 1e-9
        if sh['kind'] == 'sphere':
            assert int(sc['col_type'][c]) == COL_SPHERE and np.allclose(verts[0], sh['frame_pos'], atol=1e-9)
        else:
            assert int(sc['col_type'][c]) == COL_CAPSULE
            half = _qrot(sh['frame_quat'], [0, 0, sh['height'] / 2])
            ends = np.array([np.asarray(sh['frame_pos']) - half, np.asarray(sh['frame_pos']) + half])
            assert np.allclose(verts, ends, atol=1e-9) or np.allclose(verts, ends[::-1], atol=1e-9)

    # base: the chest
    assert np.allclose(b.bodies[hb].base_pos, mb['basePosition'], atol=1e-12) and float(sc['link_mass'][l0]) == mb['baseMass'] == 0.0
    check_shape(l0, shapes[mb['baseCollisionShapeIndex']])
    for new, old in enumerate(order):
        i, k = old - 1, l0 + 1 + new
        assert abs(float(sc['link_mass'][k]) - mb['linkMasses'][i]) < 1e-12, (new, old)
        parent_old = int(mb['linkParentIndices'][i])
        assert int(sc['link_parent'][k]) - l0 == (0 if parent_old == 0 else 1 + new_of[parent_old])
        assert np.allclose(sc['link_jpos'][k], mb['linkPositions'][i], atol=1e-12) and np.allclose(sc['link_jquat'][k], mb['linkOrientations'][i], atol=1e-12)
        assert np.allclose(sc['link_com'][k], mb['linkInertialFramePositions'][i], atol=1e-12)
        if int(mb['linkJointTypes'][i]) == 0:                        # revolute
            assert np.allclose(sc['link_axis'][k], mb['linkJointAxis'][i], atol=1e-12)
            assert abs(float(sc['link_lower'][k]) - mb['linkLowerLimits'][i]) < 1e-9 and abs(float(sc['link_upper'][k]) - mb['linkUpperLimits'][i]) < 1e-9, (new, old)
        si = int(mb['linkCollisionShapeIndices'][i])
        check_shape(k, shapes[si] if si >= 0 else None)
    # the legend of human_creation.py:5-47 is this numbering: e.g. joint 6 is the right elbow (x axis, -128 .. 0 degrees)
    assert np.allclose(sc['link_axis'][l0 + 1 + 6], [1, 0, 0]) and abs(np.rad2deg(sc['link_lower'][l0 + 1 + 6]) + 128) < 1e-6
    # self-collision filter (human_creation.py:282-299): which link pairs of the person may collide
    has_col = {k - l0 - 1 for k in set(col_link.tolist()) if l0 <= k < l0 + n}
    want = {tuple(pr) for pr in g['pairs'] if pr[0] in has_col and pr[1] in has_col and pr[0] != pr[1]}
    pl = np.asarray(sc['pair_link']).reshape(-1, 2)
    got = {(int(a) - l0 - 1, int(c) - l0 - 1) for a, c in pl if l0 <= a < l0 + n and l0 <= c < l0 + n}
    got = {(min(a, c), max(a, c)) for a, c in got}
    assert got == want
    assert abs(info['hand_radius'] - g['radii']['hand']) < 1e-12 and abs(info['elbow_radius'] - g['radii']['elbow']) < 1e-12 and abs(info['shoulder_radius'] - g['radii']['shoulder']) < 1e-12
