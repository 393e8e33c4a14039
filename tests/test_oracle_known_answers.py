"""CPU oracle vs analytic known answers (the reference has no tests, so the oracle is
pinned to physics it must reproduce exactly, not to PyBullet — parity unpinned)."""
import numpy as np
import pytest

from assistive_gym_b200 import capi
from assistive_gym_b200.scene import SceneBuilder
from oracle.oracle_py import OracleSim, gjk


def _pendulum(length=1.0, mass=1.0, axis=(1, 0, 0)):
    b = SceneBuilder()
    b.set_gravity([0, 0, -9.81])
    sh = b.create_collision_shape('sphere', radius=0.05)
    body = b.create_multibody(base_mass=0, base_pos=[0, 0, 2], link_masses=[mass], link_shapes=[sh],
                              link_positions=[[0, 0, 0]], link_orientations=[[0, 0, 0, 1]],
                              link_inertial_positions=[[0, 0, -length]], link_inertial_orientations=[[0, 0, 0, 1]],
                              link_parents=[0], link_joint_types=['revolute'], link_joint_axes=[list(axis)],
                              link_lower=[1], link_upper=[-1])
    return b.finalize(), body


def test_pendulum_period():
    sc, body = _pendulum()
    sim = OracleSim(sc, capi.default_config(dt=0.001, linear_damping=0, angular_damping=0), 1)
    sim.set_joint_state([1], q=[[0.05]])
    qs = []
    for _ in range(5000):
        sim.step(1)
        qs.append(sim.get_joint_states([1])[0][0, 0])
    zc = np.where(np.diff(np.sign(qs)) != 0)[0]
    period = 2 * np.mean(np.diff(zc)) * 0.001
    inertia = 1.0 + sc['link_inertia'][1][0]
    expected = 2 * np.pi * np.sqrt(inertia / 9.81) * (1 + 0.05 ** 2 / 16)
    assert abs(period - expected) < 2e-3


def test_pendulum_energy_drift_small():
    sc, _ = _pendulum()
    sim = OracleSim(sc, capi.default_config(dt=0.001, linear_damping=0, angular_damping=0), 1)
    sim.set_joint_state([1], q=[[0.8]])
    inertia = 1.0 + sc['link_inertia'][1][0]

    def energy():
        q, qd, _ = sim.get_joint_states([1])
        return 0.5 * inertia * qd[0, 0] ** 2 - 9.81 * np.cos(q[0, 0])
    e0 = energy()
    sim.step(3000)
    assert abs(energy() - e0) < 0.02 * abs(e0)


def test_free_fall_and_rest_on_plane():
    b = SceneBuilder()
    b.set_gravity([0, 0, -9.81])
    b.load_urdf('plane')
    sh = b.create_collision_shape('sphere', radius=0.05)
    s = b.create_multibody(base_mass=0.5, base_shape=sh, base_pos=[0, 0, 1.0])
    sc = b.finalize()
    sim = OracleSim(sc, capi.default_config(linear_damping=0, angular_damping=0), 1)
    link = [int(sc['body_link0'][s])]
    sim.step(10)   # symplectic Euler: z_n = z0 - g dt^2 n(n+1)/2
    z = sim.get_link_states(link)['pos'][0, 0, 2]
    assert abs(z - (1.0 - 9.81 * 0.02 ** 2 * 10 * 11 / 2)) < 1e-9
    sim.step(200)
    st = sim.get_link_states(link)
    assert abs(st['pos'][0, 0, 2] - 0.05) < 2e-4          # rests on the plane
    assert np.abs(st['lin_vel']).max() < 1e-3
    f = sim.contact_force_sum(s)[0]
    assert abs(f - 0.5 * 9.81) < 0.02 * 0.5 * 9.81         # contact force balances the weight


def test_coulomb_friction_threshold():
    """Box pushed sideways by gravity tilted below / above the friction cone."""
    for tilt_deg, slides in ((10.0, False), (35.0, True)):
        b = SceneBuilder()
        t = np.deg2rad(tilt_deg)
        b.set_gravity([9.81 * np.sin(t), 0, -9.81 * np.cos(t)])
        plane = b.load_urdf('plane')
        b.change_dynamics(plane, -1, lateral_friction=0.5)
        sh = b.create_collision_shape('box', half_extents=[0.1, 0.1, 0.05])
        box = b.create_multibody(base_mass=1.0, base_shape=sh, base_pos=[0, 0, 0.0505])
        b.change_dynamics(box, -1, lateral_friction=1.0)      # combined mu = 0.5 -> cone angle 26.6 deg
        sc = b.finalize()
        sim = OracleSim(sc, capi.default_config(linear_damping=0, angular_damping=0, residual_threshold=0), 1)
        sim.step(50)
        vx = sim.get_link_states([int(sc['body_link0'][box])])['lin_vel'][0, 0, 0]
        if slides:
            expect = 9.81 * (np.sin(t) - 0.5 * np.cos(t))      # a = g (sin - mu cos)
            assert abs(vx - expect * 1.0) < 0.1 * expect
        else:
            assert abs(vx) < 1e-3


def test_motor_position_control_converges_and_respects_max_force():
    sc, _ = _pendulum()
    sim = OracleSim(sc, capi.default_config(linear_damping=0, angular_damping=0), 1)
    # strong motor: reaches the target
    sim.set_motor([1], 1, target=[[0.5]], kp=[0.1], kd=[1.0], max_force=[500.0])
    sim.step(300)
    assert abs(sim.get_joint_states([1])[0][0, 0] - 0.5) < 1e-3
    # weak motor: torque saturates at max_force (1 N m < m g l sin(q))
    sim2 = OracleSim(sc, capi.default_config(linear_damping=0, angular_damping=0), 1)
    sim2.set_joint_state([1], q=[[1.2]])
    sim2.set_motor([1], 1, target=[[1.2]], kp=[0.1], kd=[1.0], max_force=[1.0])
    sim2.step(1)
    assert abs(abs(sim2.get_joint_states([1])[2][0, 0]) - 1.0) < 1e-9
    # Human.strength (human.py:86,126): a per-env scale of the force limit
    sim3 = OracleSim(sc, capi.default_config(linear_damping=0, angular_damping=0), 2)
    sim3.set_joint_state([1], q=[[1.2], [1.2]])
    sim3.set_motor([1], 1, target=[[1.2], [1.2]], kp=[0.1], kd=[1.0], max_force=[1.0])
    sim3.set_motor_force_scale([1], [[1.0], [0.25]])
    sim3.step(1)
    tau = np.abs(sim3.get_joint_states([1])[2][:, 0])
    assert abs(tau[0] - 1.0) < 1e-9 and abs(tau[1] - 0.25) < 1e-9


def test_per_body_gravity_can_be_changed_at_run_time():
    """p.setGravity(..., body=) (agent.py:196-197): the pendulum hangs still with its gravity switched off and swings with it on."""
    sc, _ = _pendulum()
    sim = OracleSim(sc, capi.default_config(linear_damping=0, angular_damping=0), 1)
    sim.set_joint_state([1], q=[[0.7]])
    sim.set_body_gravity(0, [0, 0, 0])
    sim.step(20)
    assert abs(sim.get_joint_states([1])[0][0, 0] - 0.7) < 1e-9
    sim.set_body_gravity(0, [0, 0, -9.81])
    sim.step(5)
    assert sim.get_joint_states([1])[0][0, 0] < 0.7 - 1e-3


def test_joint_limit_stops_motion():
    b = SceneBuilder()
    b.set_gravity([0, 0, -9.81])
    sh = b.create_collision_shape('sphere', radius=0.05)
    b.create_multibody(base_mass=0, base_pos=[0, 0, 2], link_masses=[1.0], link_shapes=[sh], link_positions=[[0, 0, 0]],
                       link_orientations=[[0, 0, 0, 1]], link_inertial_positions=[[0, -1.0, 0]],
                       link_inertial_orientations=[[0, 0, 0, 1]], link_parents=[0], link_joint_types=['revolute'],
                       link_joint_axes=[[1, 0, 0]], link_lower=[-0.3], link_upper=[0.3])
    sc = b.finalize()
    sim = OracleSim(sc, capi.default_config(), 1)
    sim.step(200)    # gravity torque is +x: the arm swings up to the upper limit and stays there
    q, qd, _ = sim.get_joint_states([1])
    assert 0.3 - 1e-3 < q[0, 0] < 0.3 + 0.02
    assert abs(qd[0, 0]) < 1e-6


def test_fixed_constraint_carries_payload():
    b = SceneBuilder()
    b.set_gravity([0, 0, -9.81])
    sh = b.create_collision_shape('sphere', radius=0.02)
    arm = b.create_multibody(base_mass=0, base_pos=[0, 0, 1], link_masses=[1.0], link_shapes=[sh], link_positions=[[0, 0, 0]],
                             link_orientations=[[0, 0, 0, 1]], link_inertial_positions=[[0.3, 0, 0]],
                             link_inertial_orientations=[[0, 0, 0, 1]], link_parents=[0], link_joint_types=['revolute'],
                             link_joint_axes=[[0, 1, 0]], link_lower=[1], link_upper=[-1])
    load = b.create_multibody(base_mass=0.2, base_shape=sh, base_pos=[0.5, 0, 1])
    b.set_collision_filter_pair(arm, load, 0, -1, False)
    b.create_fixed_constraint(arm, 0, load, -1, [0.2, 0, 0], [0, 0, 0], [0, 0, 0, 1], [0, 0, 0, 1], max_force=500)
    sc = b.finalize()
    sim = OracleSim(sc, capi.default_config(residual_threshold=0), 1)
    sim.set_motor([1], 1, target=[[0.0]], kp=[0.3], kd=[1.0], max_force=[100.0])
    sim.step(150)
    p = sim.get_link_states([int(sc['body_link0'][load])])['pos'][0, 0]
    assert np.linalg.norm(p - np.array([0.5, 0, 1.0])) < 5e-3     # payload stays welded to the link
    tau = sim.get_joint_states([1])[2][0, 0]
    assert abs(abs(tau) - (1.0 * 0.3 + 0.2 * 0.5) * 9.81) < 0.05 * 3.9   # motor carries both weights


def test_gjk_against_bruteforce():
    from scipy.optimize import minimize
    rng = np.random.default_rng(3)
    for _ in range(12):
        A = rng.normal(size=(rng.integers(1, 10), 3)) * 0.05
        B = rng.normal(size=(rng.integers(4, 10), 3)) * 0.05 + np.array([0.25, 0.1, -0.1])
        nA = len(A)

        def f(x):
            d = x[:nA] @ A - x[nA:] @ B
            return d @ d
        cons = [{'type': 'eq', 'fun': lambda x: x[:nA].sum() - 1}, {'type': 'eq', 'fun': lambda x: x[nA:].sum() - 1}]
        x0 = np.concatenate([np.full(nA, 1 / nA), np.full(len(B), 1 / len(B))])
        r = minimize(f, x0, bounds=[(0, 1)] * len(x0), constraints=cons, method='SLSQP', options={'ftol': 1e-15, 'maxiter': 500})
        ov, pa, pb, d = gjk(A, B)
        assert not ov
        assert abs(d - np.sqrt(r.fun)) < 1e-6
        assert abs(np.linalg.norm(pa - pb) - d) < 1e-12


def test_gjk_overlap_detected():
    cube = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=float) * 0.1
    ov, _, _, _ = gjk(cube, cube * 0.5 + 0.03)
    assert ov


def test_mass_matrix_inverse_is_symmetric_positive(feeding):
    sim = OracleSim(feeding.scene, capi.default_config(), 1)
    sim.set_joint_state(feeding.arm_links, q=np.linspace(0.5, 3.0, 7)[None])
    Mi = sim.mass_matrix_inv(feeding.robot)
    assert Mi.shape == (10, 10)
    assert np.allclose(Mi, Mi.T, atol=1e-9)
    assert np.all(np.linalg.eigvalsh(Mi) > 0)


def test_aba_matches_crba_inverse(feeding):
    """qdd from the oracle's ABA equals M^-1 tau for a pure joint-torque (damping) load."""
    cfg = capi.default_config(linear_damping=0, angular_damping=0, dt=1e-4)
    sim = OracleSim(feeding.scene, cfg, 1)
    q = np.linspace(0.5, 3.0, 7)[None]
    sim.set_joint_state(feeding.arm_links, q=q, qd=np.zeros((1, 7)))
    # zero velocity + robot gravity off -> the only generalized force is none: qdd must be ~0
    sim.set_motor(feeding.arm_links, 0, target=q, kp=[0] * 7, kd=[0] * 7, max_force=[0] * 7)
    sim.set_motor(feeding.gripper_links, 0, target=np.zeros((1, 3)), kp=[0] * 3, kd=[0] * 3, max_force=[0] * 3)
    for b in (feeding.tool, feeding.bowl, *feeding.foods):
        sim.set_body_active(b, 0)
    sim.step(1)
    qd = sim.get_joint_states(feeding.arm_links)[1]
    assert np.abs(qd).max() < 1e-10


def test_sliding_box_decelerates_at_mu_g():
    """Box launched along a horizontal plane: Coulomb friction decelerates it at mu g until it stops; the stopping
    distance is v0^2 / (2 mu g) (no damping, explicit-velocity stepping adds at most one step of v0 dt)."""
    b = SceneBuilder()
    b.set_gravity([0, 0, -9.81])
    plane = b.load_urdf('plane')
    b.change_dynamics(plane, -1, lateral_friction=0.5)
    sh = b.create_collision_shape('box', half_extents=[0.1, 0.1, 0.05])
    box = b.create_multibody(base_mass=2.0, base_shape=sh, base_pos=[0, 0, 0.0502])
    b.change_dynamics(box, -1, lateral_friction=0.8)           # combined mu = 0.4
    sc = b.finalize()
    sim = OracleSim(sc, capi.default_config(linear_damping=0, angular_damping=0, residual_threshold=0), 1)
    sim.step(10)                                                # settle the contact
    v0, mu = 1.0, 0.4
    sim.set_base_velocity(box, [[v0, 0, 0]], [[0, 0, 0]])
    link = [int(sc['body_link0'][box])]
    x0 = sim.get_link_states(link)['pos'][0, 0, 0]
    sim.step(5)
    v5 = sim.get_link_states(link)['lin_vel'][0, 0, 0]
    assert abs(v5 - (v0 - mu * 9.81 * 5 * 0.02)) < 0.02         # constant deceleration mu g
    sim.step(60)
    st = sim.get_link_states(link)
    assert abs(st['lin_vel'][0, 0, 0]) < 1e-3
    d = st['pos'][0, 0, 0] - x0
    assert abs(d - v0 ** 2 / (2 * mu * 9.81)) < 0.03, d


def test_hard_limit_clamps_and_zeroes_velocity():
    """Human.enforce_joint_limits (agent.py:240-250) as a flag: after every step a joint beyond its limit is put back on
    the limit with zero velocity -- unlike the limit ROW, which only pushes back with erp."""
    b = SceneBuilder()
    b.set_gravity([0, 0, 0])
    sh = b.create_collision_shape('sphere', radius=0.05)
    b.create_multibody(base_mass=0, base_pos=[0, 0, 2], link_masses=[1.0], link_shapes=[sh], link_positions=[[0, 0, 0]],
                       link_orientations=[[0, 0, 0, 1]], link_inertial_positions=[[0, -1.0, 0]],
                       link_inertial_orientations=[[0, 0, 0, 1]], link_parents=[0], link_joint_types=['revolute'],
                       link_joint_axes=[[1, 0, 0]], link_lower=[-0.3], link_upper=[0.3])
    sc = b.finalize()
    for hard in (False, True):
        sim = OracleSim(sc, capi.default_config(linear_damping=0, angular_damping=0), 1)
        if hard:
            sim.set_hard_limits([1], True)
        sim.set_joint_state([1], q=[[0.29]], qd=[[3.0]])       # 0.06 rad per step: crosses the limit in the first step
        sim.step(1)
        q, qd, _ = sim.get_joint_states([1])
        if hard:
            assert q[0, 0] == 0.3 and qd[0, 0] == 0.0
        else:
            assert q[0, 0] > 0.3                                # the row acts from the next step on: overshoot remains


def test_velocity_motor_tracks_target_speed():
    """VELOCITY_CONTROL row (mode 2): with ample maxForce the joint runs at the target speed after one step."""
    sc, _ = _pendulum()
    sim = OracleSim(sc, capi.default_config(linear_damping=0, angular_damping=0), 1)
    sim.set_motor([1], 2, target=[[0.7]], kp=[0.0], kd=[1.0], max_force=[500.0])
    sim.step(3)
    assert abs(sim.get_joint_states([1])[1][0, 0] - 0.7) < 1e-6
