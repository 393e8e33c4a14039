"""The oracle's articulated-body dynamics (ABA) against equations of motion derived independently:
Lagrange's equations of a three-link spatial chain, M(q) from textbook geometric Jacobians of the SCENE ARRAYS (joint frames, axes,
centre-of-mass offsets, inertia tensors), the velocity-product terms from dM/dq by central differences -- no Featherstone recursion,
no shared code.  Checks, at random states: link poses (kinematics convention), then the joint accelerations the oracle applies in
one step (mass matrix, Coriolis / centrifugal / gyroscopic terms, gravity).  (A fully symbolic sympy Lagrangian of the same chain
agrees to 1e-8 as well; it takes 13 minutes to differentiate, so it is not part of the suite.)"""
import numpy as np
import pytest

from assistive_gym_b200 import capi
from assistive_gym_b200.scene import SceneBuilder
from oracle.oracle_py import OracleSim

GRAVITY = [0.3, -0.2, -9.81]


def _chain():
    rng = np.random.default_rng(5)

    def rq():
        q = rng.normal(size=4)
        return list(q / np.linalg.norm(q))
    b = SceneBuilder()
    b.set_gravity(GRAVITY)
    shapes = [b.create_collision_shape('box', half_extents=h) for h in ([0.05, 0.12, 0.2], [0.15, 0.04, 0.08], [0.06, 0.2, 0.03])]
    axes = [list(a / np.linalg.norm(a)) for a in rng.normal(size=(3, 3))]
    body = b.create_multibody(base_mass=0, base_pos=[0.1, -0.2, 1.5], base_quat=rq(), link_masses=[1.3, 0.7, 2.1], link_shapes=shapes,
                              link_positions=[[0.1, 0.05, -0.2], [0.0, 0.3, 0.1], [-0.2, 0.1, 0.15]], link_orientations=[rq(), rq(), rq()],
                              link_inertial_positions=[[0.02, -0.1, 0.05], [0.1, 0.0, -0.07], [-0.05, 0.08, 0.1]],
                              link_inertial_orientations=[rq(), rq(), rq()], link_parents=[0, 1, 2],
                              link_joint_types=['revolute'] * 3, link_joint_axes=axes, link_lower=[1] * 3, link_upper=[-1] * 3)
    return b.finalize(), body


def _qmat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _axis_rot(a, th):
    a = np.asarray(a, dtype=np.float64)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def _lagrange(sc, links):
    """callables: link poses [(R, p)] and q-double-dot of the chain, from the scene arrays alone"""
    n = len(links)
    base = links[0] - 1
    body = int(sc['link_body'][base])
    g = np.array(GRAVITY)

    def frames(q):
        R, p = _qmat(sc['base_quat0'][body]), np.array(sc['base_pos0'][body], dtype=np.float64)
        out = []
        for i, k in enumerate(links):
            assert int(sc['link_parent'][k]) == (base if i == 0 else links[i - 1])
            p = p + R @ np.asarray(sc['link_jpos'][k], dtype=np.float64)
            Rj = R @ _qmat(sc['link_jquat'][k])
            axis_w = Rj @ np.asarray(sc['link_axis'][k], dtype=np.float64)
            R = Rj @ _axis_rot(sc['link_axis'][k], q[i])
            out.append((R, p.copy(), axis_w))
        return out

    def mass_matrix_and_potential_gradient(q):
        fr = frames(q)
        M, dV = np.zeros((n, n)), np.zeros(n)
        for i, k in enumerate(links):
            R, p, _ = fr[i]
            com = p + R @ np.asarray(sc['link_com'][k], dtype=np.float64)
            Jv, Jw = np.zeros((3, n)), np.zeros((3, n))
            for j in range(i + 1):                       # joint j moves link i: v_com = a_j x (com - o_j) qd_j, omega = a_j qd_j
                Jv[:, j] = np.cross(fr[j][2], com - fr[j][1])
                Jw[:, j] = fr[j][2]
            Ri = R @ _qmat(sc['link_iquat'][k])
            Iw = Ri @ np.diag(np.asarray(sc['link_inertia'][k], dtype=np.float64)) @ Ri.T
            m = float(sc['link_mass'][k])
            M += m * Jv.T @ Jv + Jw.T @ Iw @ Jw
            dV += -m * (g @ Jv)                          # V = -m g.com
        return M, dV

    def qdd(q, qd, h=1e-6):
        M, dV = mass_matrix_and_potential_gradient(q)
        dM = []
        for i in range(n):
            e = np.zeros(n); e[i] = h
            dM.append((mass_matrix_and_potential_gradient(q + e)[0] - mass_matrix_and_potential_gradient(q - e)[0]) / (2 * h))
        Mdot = sum(dM[i] * qd[i] for i in range(n))
        dT = np.array([0.5 * qd @ dM[i] @ qd for i in range(n)])
        # d/dt (M qd) - dT/dq + dV/dq = 0
        return np.linalg.solve(M, dT - dV - Mdot @ qd)

    def poses(q):
        return [np.hstack([R, p[:, None]]) for R, p, _ in frames(q)]
    return poses, qdd


@pytest.fixture(scope='module')
def chain():
    sc, body = _chain()
    l0 = int(sc['body_link0'][body])
    links = [l0 + 1, l0 + 2, l0 + 3]
    return sc, links, _lagrange(sc, links)


def test_link_poses_match_the_independent_kinematics(chain):
    sc, links, (fposes, _) = chain
    sim = OracleSim(sc, capi.default_config(), 4)
    rng = np.random.default_rng(0)
    q = rng.uniform(-2.5, 2.5, size=(4, 3))
    sim.set_joint_state(links, q=q, qd=np.zeros_like(q))
    sim.forward_kinematics()
    st = sim.get_link_states(links)
    for e in range(4):
        for i, Rp in enumerate(fposes(q[e])):
            Rp = np.array(Rp, dtype=np.float64)
            assert np.allclose(Rp[:, 3], st['pos'][e, i], atol=1e-12)
            assert np.allclose(Rp[:, :3], _qmat(st['quat'][e, i]), atol=1e-12)


def test_joint_accelerations_match_the_lagrangian(chain):
    sc, links, (_, qdd) = chain
    dt = 1e-3
    n = 6
    sim = OracleSim(sc, capi.default_config(dt=dt, linear_damping=0, angular_damping=0), n)
    rng = np.random.default_rng(1)
    q = rng.uniform(-2.5, 2.5, size=(n, 3))
    qd = rng.uniform(-3, 3, size=(n, 3))
    qd[0] = 0                                   # gravity and the mass matrix alone
    sim.set_joint_state(links, q=q, qd=qd)
    sim.step(1)
    q1, qd1, _ = sim.get_joint_states(links)
    for e in range(n):
        want = qdd(q[e], qd[e])
        got = (qd1[e] - qd[e]) / dt
        assert np.allclose(got, want, rtol=1e-6, atol=1e-6), (e, got, want)
        assert np.allclose(q1[e], q[e] + dt * qd1[e], atol=1e-12)      # symplectic Euler: the new velocity moves the joints


def test_product_kernel_bodies_match_the_lagrangian(chain, emu_lib):
    """the same check on the product's K1 (`dyn_body`, fp32; kernel bodies compiled for the host): no oracle involved"""
    from assistive_gym_b200.sim import BatchSim
    sc, links, (_, qdd) = chain
    dt = 0.02
    n = 6
    sim = BatchSim(sc, capi.default_config(dt=dt, linear_damping=0, angular_damping=0), n, _lib=emu_lib)
    rng = np.random.default_rng(2)
    q = rng.uniform(-2.5, 2.5, size=(n, 3))
    qd = rng.uniform(-3, 3, size=(n, 3))
    sim.set_joint_state(links, q=q, qd=qd)
    sim.step(1)
    _, qd1, _ = sim.get_joint_states(links)
    for e in range(n):
        want = qdd(q[e].astype(np.float32).astype(np.float64), qd[e].astype(np.float32).astype(np.float64))
        got = (qd1[e] - qd[e].astype(np.float32)) / dt
        assert np.allclose(got, want, rtol=2e-3, atol=2e-3), (e, got, want)
