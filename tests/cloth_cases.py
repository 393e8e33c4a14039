"""Shared cloth parity cases: the product (CUDA build, or the host-compiled kernel bodies) next to the CPU oracle.

Small synthetic scenes exercise each piece of the cloth step (free fall + drag, links, anchors, contacts with a sphere /
capsule / box / plane, a MOVING articulated collider) and one case uses the reference's gown mesh and parameters.
"""
import numpy as np

from assistive_gym_b200 import capi
from assistive_gym_b200.cloth import ClothModel
from assistive_gym_b200.scene import SceneBuilder, quat_from_rpy


def grid_cloth(nx=24, ny=16, spacing=0.02, params=None):
    """A rectangular sheet of nx x ny nodes triangulated with alternating diagonals."""
    idx = lambda i, j: i * ny + j
    verts = np.array([[i * spacing, j * spacing, 0.0] for i in range(nx) for j in range(ny)], dtype=np.float64)
    faces = []
    for i in range(nx - 1):
        for j in range(ny - 1):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            faces += [[a, b, c], [a, c, d]] if (i + j) % 2 == 0 else [[a, b, d], [b, c, d]]
    p = dict(total_mass=0.16 * len(verts) / 3966.0)
    if params:
        p.update(params)
    return ClothModel(verts, np.array(faces, dtype=np.int32), scale=1.0, params=p)


def obstacle_scene():
    """plane + a static multibody (sphere base, capsule and box on fixed links) + a 1-DoF arm swinging a capsule.
    Returns (scene arrays, collider links, static flags, moving joint link)."""
    b = SceneBuilder()
    b.world_gravity = np.array([0.0, 0.0, -9.81])
    plane = b.load_urdf('plane')
    sph = b.create_collision_shape('sphere', radius=0.08)
    cap = b.create_collision_shape('capsule', radius=0.04, height=0.3, frame_quat=quat_from_rpy([0, np.pi / 2, 0]))
    box = b.create_collision_shape('box', half_extents=(0.06, 0.05, 0.04))
    stat = b.create_multibody(base_mass=0.0, base_shape=sph, base_pos=(0.1, 0.1, 0.3),
                              link_masses=[0.0, 0.0], link_shapes=[cap, box], link_positions=[(0.25, 0.05, 0.0), (0.1, 0.25, -0.05)],
                              link_orientations=[(0, 0, 0, 1)] * 2, link_inertial_positions=[(0, 0, 0)] * 2,
                              link_inertial_orientations=[(0, 0, 0, 1)] * 2, link_parents=[0, 0],
                              link_joint_types=['fixed', 'fixed'], link_joint_axes=[(0, 0, 1)] * 2, name='obstacles')
    arm_c = b.create_collision_shape('capsule', radius=0.03, height=0.25, frame_pos=(0.0, 0.0, 0.125))
    arm = b.create_multibody(base_mass=0.0, base_shape=-1, base_pos=(0.3, 0.2, 0.12),
                             link_masses=[1.0], link_shapes=[arm_c], link_positions=[(0, 0, 0)], link_orientations=[(0, 0, 0, 1)],
                             link_inertial_positions=[(0, 0, 0.125)], link_inertial_orientations=[(0, 0, 0, 1)], link_parents=[0],
                             link_joint_types=['revolute'], link_joint_axes=[(0, 1, 0)], link_lower=[-3.0], link_upper=[3.0], name='arm')
    b.set_gravity([0, 0, 0], body=arm)
    scene = b.finalize()
    gl = lambda body, k: int(scene['body_link0'][body]) + 1 + k
    links = [gl(plane, -1), gl(stat, -1), gl(stat, 0), gl(stat, 1), gl(arm, 0)]
    static = [1, 1, 1, 1, 0]
    return scene, links, static, gl(arm, 0)


def make_pair(make_product, make_oracle, model, n=2, substeps=8, col=True, anchors=(0, 5), gravity=(0, 0, -9.81), seed=0,
              height=0.45, arm_speed=2.0):
    """Two simulations in the same state: the sheet hovering over the obstacles, slightly crumpled, anchors held."""
    scene, links, static, arm_joint = obstacle_scene()
    cfg = capi.default_config(num_substeps=substeps)
    sims = [make_product(scene, cfg, n), make_oracle(scene, cfg, n)]
    rng = np.random.default_rng(seed)
    x0 = np.repeat(model.rest[None], n, axis=0) + np.array([0.0, 0.0, height])
    x0 = x0 + rng.normal(scale=1e-3, size=x0.shape)
    x0[:, :, 0] += rng.uniform(-0.02, 0.02, size=(n, 1))
    v0 = rng.normal(scale=0.05, size=x0.shape)
    anchor_pos = x0[:, anchors[0]].copy() if len(anchors) else np.zeros((n, 3))
    local = (model.rest[list(anchors)] - model.rest[anchors[0]]) if len(anchors) else np.zeros((0, 3))
    for s in sims:
        s.cloth_init(model, links if col else [], static if col else [], list(anchors), local, gravity=gravity)
        s.cloth_set_state(x0, v0)
        s.cloth_set_anchor(anchor_pos)
        s.set_joint_state([arm_joint], q=np.full((n, 1), -0.8), qd=np.full((n, 1), arm_speed))
        s.forward_kinematics()
    return sims, scene, arm_joint


def compare(sims, steps):
    prod, orc = sims
    prod.step(steps)
    orc.step(steps)
    xp, vp = prod.cloth_get_state()
    xo, vo = orc.cloth_get_state()
    cp = prod.cloth_get_contacts(2048)
    co = orc.cloth_get_contacts(2048)
    return dict(dx=float(np.abs(xp - xo).max()), dv=float(np.abs(vp - vo).max()), travel=float(np.abs(xo).max()),
                contacts_prod=cp, contacts_orc=co, x=xo, v=vo, xp=xp)


def contact_sets_equal(cp, co, force_rtol=2e-2, force_atol=1e-3):
    """Same (node, link) sets per env; forces agree."""
    worst = 0.0
    for e in range(len(cp[0])):
        kp = {(int(cp[1][e, k]), int(cp[4][e, k])): cp[3][e, k] for k in range(cp[0][e])}
        ko = {(int(co[1][e, k]), int(co[4][e, k])): co[3][e, k] for k in range(co[0][e])}
        if set(kp) != set(ko):
            return False, 'env %d: contact sets differ by %d' % (e, len(set(kp) ^ set(ko)))
        for key in kp:
            err = np.abs(kp[key] - ko[key]).max()
            ref = np.abs(ko[key]).max()
            if err > force_atol + force_rtol * ref:
                return False, 'env %d contact %s: force %s vs %s' % (e, key, kp[key], ko[key])
            worst = max(worst, err)
    return True, worst
