"""The repo's restatements of the reference's geometric helpers against vectors recorded from the reference's own code
(tests/golden/util_vectors.npz, written by tests/golden/make_golden_util.py from assistive_gym/envs/util.py):
  * `bed_bathing_batch.capsule_points` (util.py:80-113): the wiping targets the BedBathing kernels are fed with;
  * `tests/dressing_cases.line_intersects_triangle` / `sleeve_on_arm_reward` (util.py:125-202): the numpy reference the fused
    Dressing kernel (`dressing_post_body`) is tested against in tests/test_dressing.py -- so the chain reference -> restatement ->
    kernel is closed on both ends."""
import os

import numpy as np
import pytest

from assistive_gym_b200.bed_bathing_batch import capsule_points
from tests.dressing_cases import line_intersects_triangle, sleeve_on_arm_reward

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'util_vectors.npz'))


def test_capsule_points_are_the_reference_s():
    off = 0
    for row, n in zip(G['capsule_in'], G['capsule_n']):
        want = G['capsule_pts'][off:off + n]; off += n
        got = capsule_points(row[0:3], row[3:6], row[6], row[7])
        assert got.shape == want.shape, row
        assert np.allclose(got, want, atol=1e-12), row
    assert off == len(G['capsule_pts'])


def test_line_intersects_triangle_is_the_reference_s():
    got = np.array([bool(line_intersects_triangle(r[0:3], r[3:6], r[6:9], r[9:12], r[12:15])) for r in G['tri_in']])
    assert np.array_equal(got, G['tri_out']) and 0 < got.sum() < len(got)


def test_sleeve_on_arm_reward_is_the_reference_s():
    want = G['sleeve_out']
    assert want[:, 0].sum() > 20 and want[:, 1].sum() > 10                     # both "in sleeve" outcomes occur among the cases
    for r, w in zip(G['sleeve_in'], want):
        t1, t2 = r[0:9].reshape(3, 3), r[9:18].reshape(3, 3)
        got = sleeve_on_arm_reward(t1, t2, r[18:21], r[21:24], r[24:27], r[27], r[28], r[29])
        # reference returns (forearm, upperarm, along forearm, along upperarm, to hand, to elbow, to shoulder, forearm length, upperarm length)
        assert bool(got[0]) == bool(w[0]) and bool(got[1]) == bool(w[1])
        assert np.allclose([got[2], got[3], got[4], got[5], got[6]], [w[2], w[3], w[4], w[7], w[8]], atol=1e-12)
