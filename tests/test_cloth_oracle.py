"""Analytic known answers for the cloth ORACLE (oracle/oracle_cloth.h).  The oracle restates Bullet's btSoftBody solver
from memory (parity unpinned against Bullet); these tests pin it to closed forms of the algorithm it states."""
import numpy as np
import pytest

from assistive_gym_b200 import capi
from assistive_gym_b200.scene import SceneBuilder
from oracle.oracle_py import OracleSim
from tests.cloth_cases import grid_cloth

DT = 0.0025


def _plane_scene():
    b = SceneBuilder()
    b.world_gravity = np.array([0.0, 0.0, -9.81])
    plane = b.load_urdf('plane')
    sc = b.finalize()
    return sc, int(sc['body_link0'][plane])


def _sim(model, x0, v0=None, anchors=(), gravity=(0, 0, -9.81), plane=False, friction=None):
    sc, pl = _plane_scene()
    s = OracleSim(sc, capi.default_config(dt=DT, num_substeps=1), 1)
    s.cloth_init(model, [pl] if plane else [], [1] if plane else [], list(anchors), model.rest[list(anchors)] - (model.rest[anchors[0]] if anchors else 0), gravity=gravity)
    s.cloth_set_state(x0[None], (np.zeros_like(x0) if v0 is None else v0)[None])
    if anchors:
        s.cloth_set_anchor(x0[anchors[0]][None])
    if friction is not None:
        s.set_link_friction(pl, np.array([friction]))
    return s


def test_free_fall_is_symplectic_euler():
    m = grid_cloth(params=dict(kDG=0.0, kDP=0.0))
    x0 = m.rest + np.array([0, 0, 5.0])
    s = _sim(m, x0)
    n = 40
    s.step(n)
    x, v = s.cloth_get_state()
    assert np.allclose(v[0, :, 2], -9.81 * DT * n, rtol=1e-9)
    assert np.allclose(x[0, :, 2], 5.0 - 9.81 * DT * DT * n * (n + 1) / 2, rtol=1e-9)
    assert np.abs(x[0, :, :2] - x0[:, :2]).max() < 1e-12               # unstretched links do nothing


def test_kdp_damps_velocity_geometrically():
    m = grid_cloth(params=dict(kDG=0.0, kDP=0.01))
    v0 = np.tile([0.3, -0.2, 0.1], (m.n_nodes, 1))
    s = _sim(m, m.rest + np.array([0, 0, 5.0]), v0=v0, gravity=(0, 0, 0))
    s.step(10)
    _, v = s.cloth_get_state()
    assert np.allclose(v[0], v0 * 0.99 ** 10, rtol=1e-9)


def test_drag_balances_weight_at_the_terminal_velocity():
    """addAeroForceToNode, V_Point: a node moving along its normal feels kDG * rho * area * |v|^3 / 2 against the motion.
    A flat horizontal sheet (uniform diagonals: every interior node owns area s^2) started at the speed where that balances
    the weight keeps it: one step leaves the interior velocities unchanged to first order, and equals the closed form."""
    from assistive_gym_b200.cloth import ClothModel
    nx = ny = 12
    sp = 0.02
    idx = lambda i, j: i * ny + j
    verts = np.array([[i * sp, j * sp, 0.0] for i in range(nx) for j in range(ny)])
    faces = []
    for i in range(nx - 1):
        for j in range(ny - 1):
            faces += [[idx(i, j), idx(i + 1, j), idx(i + 1, j + 1)], [idx(i, j), idx(i + 1, j + 1), idx(i, j + 1)]]
    m = ClothModel(verts, np.array(faces), params=dict(kDG=10.0, kDP=0.0, total_mass=0.16 * nx * ny / 3966.0))
    mass, area = 1.0 / m.inv_mass, sp * sp
    vt = np.cbrt(2 * mass * 9.81 / (area * 1.2 * 10.0))
    v0 = np.tile([0.0, 0.0, -vt], (m.n_nodes, 1))
    s = _sim(m, m.rest + np.array([0, 0, 50.0]), v0=v0)
    s.step(1)
    _, v = s.cloth_get_state()
    inner = [idx(i, j) for i in range(4, 8) for j in range(4, 8)]        # two rings away from the boundary: neighbours move alike
    assert np.allclose(m.node_area[m.rank][inner], area)
    vp = -vt - 9.81 * DT                                                  # after gravity
    want = vp + (area * abs(vp) ** 3 / 2 * 1.2 * 10.0) * m.inv_mass * DT  # drag points up
    assert np.allclose(v[0, inner, 2], want, rtol=1e-9)
    assert abs(want + vt) < 1e-2 * vt                                     # weight and drag cancel at vt (to first order in dt)


def test_one_position_iteration_of_a_stretched_triangle():
    """PSolve_Links, sequential over the list: each link moves both ends by del * (c1 - len) / (c1 + len) * kLST / 2."""
    from assistive_gym_b200.cloth import ClothModel
    verts = np.array([[0, 0, 0], [0.1, 0, 0], [0, 0.1, 0.0]])
    m = ClothModel(verts, np.array([[0, 1, 2]]), params=dict(kDG=0.0, kDP=0.0, piterations=1, total_mass=0.003), reorder=False)
    x0 = verts * 1.5 + np.array([0, 0, 1.0])
    s = _sim(m, x0, gravity=(0, 0, 0))
    s.step(1)
    x, _ = s.cloth_get_state()
    want = x0.copy()
    for (i, j), r2 in zip(m.links, m.link_rest2):
        d = want[j] - want[i]
        k = (r2 - d @ d) / (r2 + d @ d) * 0.055 / 2
        want[i] -= d * k
        want[j] += d * k
    assert np.allclose(x[0], want, atol=1e-14)


def test_anchor_holds_a_hanging_sheet():
    m = grid_cloth(nx=10, ny=6)
    x0 = m.rest + np.array([0, 0, 2.0])
    s = _sim(m, x0, anchors=(0, 5))
    s.step(800)
    x, v = s.cloth_get_state()
    assert np.linalg.norm(x[0, 0] - x0[0]) < 1e-2 and np.linalg.norm(x[0, 5] - x0[5]) < 1e-2     # kAHR = 1: the anchors stay put (the links of the last iteration tug them by mm)
    assert x[0, :, 2].min() < 2.0 - 0.1                                  # the rest hangs below them
    # (PSolve_Anchors as restated flips the anchored node's offset every iteration: with an odd iteration count and kAHR = 1
    #  the anchored nodes keep a mm-sized period-2 jitter; everything else comes to rest)
    free = np.setdiff1d(np.arange(m.n_nodes), [0, 5])
    assert np.abs(v[0, free]).max() < 0.1


def test_sheet_rests_on_the_plane_at_the_collision_margin_and_friction_stops_it():
    m = grid_cloth(nx=10, ny=10)
    x0 = m.rest + np.array([0, 0, 0.05])
    v0 = np.tile([0.5, 0.0, 0.0], (m.n_nodes, 1))
    s = _sim(m, x0, v0=v0, plane=True, friction=1.0)
    s.step(400)
    x, v = s.cloth_get_state()
    assert np.abs(x[0, :, 2] - 0.04).max() < 1.5e-3                     # collisionMargin = 0.04 (dressing.py:146)
    assert np.abs(v[0, :, 2]).max() < 0.2                                # resting chatter at the margin: < 0.5 mm per substep
    assert np.abs(v[0, :, 0]).max() < 0.02                               # kDF * friction = 0.39: it has stopped
    assert 0.0 < x[0, :, 0].mean() - x0[:, 0].mean() < 0.5 * 400 * DT   # after sliding a bit
    # resting contact chatters (a node sitting exactly at the margin is in contact every other substep or so): the
    # time-averaged contact force carries the weight
    fz, cn = [], []
    for _ in range(80):
        s.step(1)
        cnt, node, pos, force, link = s.cloth_get_contacts(256)
        fz.append(force[0, :cnt[0], 2].sum()); cn.append(cnt[0])
    weight = 9.81 * m.n_nodes / m.inv_mass
    assert min(cn) > 10 and abs(np.mean(fz) - weight) / weight < 0.15, (min(cn), np.mean(fz), weight)
