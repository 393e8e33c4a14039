"""Capsule human body (42 links) — data-driven restatement of the body the reference builds in
`envs/human_creation.py:58-316` via `p.createCollisionShape` / `p.createMultiBody`.

The numbers (segment radii/lengths, joint offsets, mass fractions, joint axes and limits) are the
reference's model parameters (human_creation.py:72-173 dimensions, :185-278 link tables, :284-299
self-collision filter); the construction code here is table-driven and goes through SceneBuilder.
Public link/joint indices after DFS re-indexing match `envs/agents/human.py:5-58`
(right arm 0-9, left arm 10-19, head 20-23, waist 24-27, right leg 28-34, left leg 35-41).
"""
import numpy as np

from .scene import quat_from_rpy

HALF_PI = np.pi / 2.0

# per-gender dimensions: (male, female)
DIMS = {
    'mass': (78.4, 62.5),
    'chest_r': (0.127, 0.127), 'chest_l': (0.056, 0.01),
    'sh_r': (0.106, 0.092), 'sh_w': (0.253, 0.225),
    'neck_r': (0.06, 0.05), 'neck_l': (0.124, 0.121),
    'uarm_r': (0.043, 0.0355), 'uarm_l': (0.279, 0.264),
    'farm_r': (0.033, 0.027), 'farm_l': (0.257, 0.234),
    'waist_r': (0.1205, 0.11), 'waist_l': (0.049, 0.009),
    'hips_r': (0.1335, 0.127), 'hips_l': (0.094, 0.117), 'hips_z': (0.08125, 0.075),
    'thigh_r': (0.08, 0.0775), 'thigh_l': (0.424, 0.391), 'thigh_dx': (0.009, 0.0145),
    'shin_r': (0.05, 0.045), 'shin_l': (0.403, 0.367),
    'foot_r': (0.05, 0.045), 'foot_l': (0.215, 0.195), 'foot_dy': (0.1, 0.09), 'foot_dz': (0.025, 0.0225),
    'chest_z': (1.2455, 1.148), 'shoulders_z': (0.1415 / 2, 0.132 / 2), 'neck_z': (0.1515, 0.132),
    'head_z': (0.399 - 0.1415 - 0.1205, 0.12), 'uarm_dx': (0.073, 0.067), 'waist_z': (0.156, 0.15),
    'head_pos': ([0.09, 0.08, -0.07 + 0.01], [-0.089, -0.09, -0.07]),
    'head_mesh': ('head_male_vhacd', 'head_female_vhacd'),
}
HEAD_SCALE = 0.89


def _deg(*v):
    return [np.deg2rad(x) for x in v]


def create_human(b, gender='male', static=True, limit_scale=1.0, mass=None, radius_scale=1.0, height_scale=1.0, cloth=False):
    """Add the human to SceneBuilder `b`; returns (body id, info dict with hand/elbow/shoulder radius)."""
    g = 0 if gender == 'male' else 1
    D = {k: v[g] for k, v in DIMS.items()}
    m = D['mass'] if mass is None else mass
    rs, hs = radius_scale, height_scale
    qy = quat_from_rpy([0, HALF_PI, 0])
    qx = quat_from_rpy([HALF_PI, 0, 0])

    def cap(r, l, off=(0, 0, 0), quat=(0, 0, 0, 1)):
        return b.create_collision_shape('capsule', radius=r, height=l, frame_pos=off, frame_quat=quat)

    def sph(r, off=(0, 0, 0)):
        return b.create_collision_shape('sphere', radius=r, frame_pos=off)

    S = {
        'none': -1,
        'chest': cap(D['chest_r'] * rs, D['chest_l'], quat=qy),
        'r_sh': cap(D['sh_r'] * rs, D['sh_w'] / 8, off=[-D['sh_w'] / 2.5 + D['sh_w'] / 16, 0, 0], quat=qy),
        'l_sh': cap(D['sh_r'] * rs, D['sh_w'] / 8, off=[D['sh_w'] / 2.5 - D['sh_w'] / 16, 0, 0], quat=qy),
        'neck': cap(D['neck_r'] * rs, D['neck_l'] * hs, off=[0, 0, (0.2565 - 0.1415 - 0.025) * hs]),
        'uarm': cap(D['uarm_r'] * rs, D['uarm_l'] * hs, off=[0, 0, -D['uarm_l'] / 2.0 * hs]),
        'farm': cap(D['farm_r'] * rs, D['farm_l'] * hs, off=[0, 0, -D['farm_l'] / 2.0 * hs]),
        'hand': sph(D['uarm_r'] * rs, off=[0, 0, -D['uarm_r'] * rs]),
        'waist': cap(D['waist_r'] * rs, D['waist_l'], quat=qy),
        'hips': cap(D['hips_r'] * rs, D['hips_l'], off=[0, 0, -D['hips_z'] * hs], quat=qy),
        'thigh': cap(D['thigh_r'] * rs, D['thigh_l'] * hs, off=[0, 0, -D['thigh_l'] / 2.0 * hs]),
        'shin': cap(D['shin_r'] * rs, D['shin_l'] * hs, off=[0, 0, -D['shin_l'] / 2.0 * hs]),
        'foot': cap(D['foot_r'] * rs, D['foot_l'] * hs, off=[0, -D['foot_dy'], -D['foot_dz'] * rs], quat=qx),
        'head': b.create_collision_shape('mesh', mesh_asset=D['head_mesh'], mesh_scale=[HEAD_SCALE] * 3,
                                         frame_pos=D['head_pos'], frame_quat=qx),
    }
    if cloth:   # extra spheres at the arm joints so cloth cannot slip into the capsule ends (human_creation.py:96-101)
        S['sh_cloth'] = sph(D['uarm_r'] * rs)
        S['el_cloth'] = sph(D['uarm_r'] * rs)
        S['wr_cloth'] = sph(D['farm_r'] * rs)
    else:
        S['sh_cloth'] = S['el_cloth'] = S['wr_cloth'] = -1
    Z = [0.0, 0.0, 0.0]
    P = {
        'joint': Z, 'shoulders': [0, 0, D['shoulders_z'] * hs], 'neck': [0, 0, D['neck_z'] * hs], 'head': [0, 0, D['head_z'] * hs],
        'r_uarm': [-D['sh_r'] * rs - D['uarm_dx'], 0, 0], 'l_uarm': [D['sh_r'] * rs + D['uarm_dx'], 0, 0],
        'farm': [0, 0, -D['uarm_l'] * hs], 'hand': [0, 0, -(D['farm_r'] * rs + D['farm_l'] * hs)],
        'waist': [0, 0, -D['waist_z'] * hs], 'hips': [0, 0, -D['hips_z'] * hs],
        'r_thigh': [-D['thigh_r'] * rs - D['thigh_dx'], 0, -D['hips_z'] * hs], 'l_thigh': [D['thigh_r'] * rs + D['thigh_dx'], 0, -D['hips_z'] * hs],
        'shin': [0, 0, -D['thigh_l'] * hs], 'foot': [0, 0, -D['shin_l'] * hs - D['foot_dz']],
    }
    X, Y, ZA = [1, 0, 0], [0, 1, 0], [0, 0, 1]
    # (parent creation index [0 = base], mass fraction, shape, position, axis, lower deg, upper deg, scaled by limit_scale)
    rows = []

    def add(parent, frac, shape, pos, axis, lo, hi, scaled=True, fixed=False):
        rows.append((parent, frac, S[shape], P[pos], axis, lo, hi, scaled, fixed))

    # shoulders, neck, head (creation links 1-10)
    add(0, 0, 'none', 'shoulders', X, -10, 10); add(1, 0, 'none', 'shoulders', Y, -10, 30); add(2, .05, 'r_sh', 'joint', ZA, -35, 35)
    add(0, 0, 'none', 'shoulders', X, -10, 10); add(4, 0, 'none', 'shoulders', Y, -30, 10); add(5, .05, 'l_sh', 'joint', ZA, -35, 35)
    add(0, .01, 'neck', 'neck', X, -10, 20); add(7, 0, 'none', 'head', X, -50, 50); add(8, 0, 'none', 'joint', Y, -34, 34); add(9, .07, 'head', 'joint', ZA, -70, 70)
    # right arm (11-17), left arm (18-24)
    arm_shapes = ['none', 'sh_cloth', 'uarm', 'el_cloth', 'farm', 'wr_cloth', 'hand']
    arm_axes = [Y, X, ZA, X, ZA, X, Y]
    arm_frac = [0, 0, .033, 0, .019, 0, .0065]
    r_lim = [(5, 198), (-188, 61), (-90, 90), (-128, 0), (-90, 90), (-81, 90), (-27, 47)]
    l_lim = [(-198, -5), (-188, 61), (-90, 90), (-128, 0), (-90, 90), (-81, 90), (-47, 27)]
    for side, first_parent, lims, pos0 in (('r', 3, r_lim, 'r_uarm'), ('l', 6, l_lim, 'l_uarm')):
        base = len(rows)
        poss = [pos0, 'joint', 'joint', 'farm', 'joint', 'hand', 'joint']
        for i in range(7):
            add(first_parent if i == 0 else base + i, arm_frac[i], arm_shapes[i], poss[i], arm_axes[i], lims[i][0], lims[i][1])
    # waist and hips (25-28)
    add(0, 0, 'waist', 'waist', [0, 0, 0], 0, 0, scaled=False, fixed=True)
    add(25, 0, 'none', 'hips', X, -75, 30, scaled=False); add(26, .13, 'none', 'joint', Y, -30, 30, scaled=False); add(27, .14, 'hips', 'joint', ZA, -30, 30, scaled=False)
    # legs (29-35, 36-42)
    leg_shapes = ['none', 'none', 'thigh', 'shin', 'none', 'none', 'foot']
    leg_axes = [X, Y, ZA, X, X, Y, ZA]
    leg_frac = [0, 0, .105, .0475, 0, 0, .014]
    r_leg = [(-127, 30), (-40, 45), (-45, 40), (0, 130), (-35, 38), (-23, 24), (-43, 35)]
    l_leg = [(-127, 30), (-45, 40), (-40, 45), (0, 130), (-35, 38), (-24, 23), (-35, 43)]
    for lims, pos0 in ((r_leg, 'r_thigh'), (l_leg, 'l_thigh')):
        base = len(rows)
        poss = [pos0, 'joint', 'joint', 'shin', 'foot', 'joint', 'joint']
        for i in range(7):
            add(28 if i == 0 else base + i, leg_frac[i], leg_shapes[i], poss[i], leg_axes[i], lims[i][0], lims[i][1], scaled=False)
    n = len(rows)
    assert n == 42
    # NOTE: the reference lists the mass-bearing waist links as [0, 0, .13, .14] over (waist, hips_x, hips_y, hips_z)
    body = b.create_multibody(
        base_mass=0 if static else m * 0.1, base_shape=S['chest'], base_pos=[0, 0, D['chest_z'] * hs], base_quat=[0, 0, 0, 1],
        link_masses=[m * r[1] for r in rows], link_shapes=[r[2] for r in rows], link_positions=[r[3] for r in rows],
        link_orientations=[[0, 0, 0, 1]] * n, link_inertial_positions=[[0, 0, 0]] * n, link_inertial_orientations=[[0, 0, 0, 1]] * n,
        link_parents=[r[0] for r in rows], link_joint_types=['fixed' if r[8] else 'revolute' for r in rows],
        link_joint_axes=[r[4] for r in rows],
        link_lower=[np.deg2rad(r[5]) * (limit_scale if r[7] else 1.0) for r in rows],
        link_upper=[np.deg2rad(r[6]) * (limit_scale if r[7] else 1.0) for r in rows],
        self_collision=True, name='human_' + gender)
    # self collision: everything off, then arms / legs against the rest (human_creation.py:284-299)
    nj = b.num_joints(body)
    for i in range(-1, nj):
        for j in range(-1, nj):
            b.set_collision_filter_pair(body, body, i, j, False)
    for i in range(3, 10):
        for j in [-1] + list(range(10, nj)):
            b.set_collision_filter_pair(body, body, i, j, True)
    for i in range(13, 20):
        for j in list(range(-1, 10)) + list(range(20, nj)):
            b.set_collision_filter_pair(body, body, i, j, True)
    for i in range(28, 35):
        for j in list(range(-1, 24)) + list(range(35, nj)):
            b.set_collision_filter_pair(body, body, i, j, True)
    for i in range(35, nj):
        for j in list(range(-1, 24)) + list(range(28, 35)):
            b.set_collision_filter_pair(body, body, i, j, True)
    info = {'hand_radius': D['uarm_r'] * rs, 'elbow_radius': D['uarm_r'] * rs, 'shoulder_radius': D['uarm_r'] * rs,
            'mass': m, 'gender': gender}
    return body, info
