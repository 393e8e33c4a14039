"""ctypes wrapper of the CPU oracle (oracle/agphys_oracle.cpp).

TEST INFRASTRUCTURE — PARITY UNPINNED (the reference path lives in PyBullet, which is not
available here; see oracle/README.md).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)


def build(force=False):
    so = os.path.join(_HERE, '_build', 'liboracle.so')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, 'agphys_oracle.cpp')):
        subprocess.check_call(['make', '-s', '-C', _HERE], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


_libs = {}


def _load(f32=False):
    key = bool(f32)
    if key in _libs:
        return _libs[key]
    import sys
    sys.path.insert(0, _REPO)
    from assistive_gym_b200.capi import AgConfig, AgSceneDesc
    build()
    lib = C.CDLL(os.path.join(_HERE, '_build', 'liboracle_f32.so' if f32 else 'liboracle.so'))
    vp, ci = C.c_void_p, C.c_int
    lib.oracle_create.restype = vp
    lib.oracle_create.argtypes = [C.POINTER(AgSceneDesc), C.POINTER(AgConfig), ci]
    lib.oracle_destroy.argtypes = [vp]
    lib.oracle_num_dofs.argtypes = [vp, ci]
    lib.oracle_set_base_pose.argtypes = [vp, ci, vp, vp, vp]
    lib.oracle_set_base_velocity.argtypes = [vp, ci, vp, vp, vp]
    lib.oracle_set_joint_state.argtypes = [vp, ci, vp, vp, vp, vp]
    lib.oracle_set_link_friction.argtypes = [vp, ci, vp, vp]
    lib.oracle_set_body_mode.argtypes = [vp, ci, vp]
    lib.oracle_set_motor.argtypes = [vp, ci, vp, ci, vp, vp, vp, vp]
    lib.oracle_set_hard_limits.argtypes = [vp, ci, vp, ci]
    lib.oracle_set_motor_targets.argtypes = [vp, ci, vp, vp]
    lib.oracle_set_motor_force_scale.argtypes = [vp, ci, vp, vp]
    lib.oracle_forward_kinematics.argtypes = [vp]
    lib.oracle_set_body_gravity.argtypes = [vp, ci, vp]
    lib.oracle_step.argtypes = [vp, ci, ci]
    lib.oracle_get_joint_states.argtypes = [vp, ci, vp, vp, vp, vp]
    lib.oracle_get_link_states.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp, vp]
    lib.oracle_get_contacts.argtypes = [vp, ci, ci, ci, ci, ci, vp, vp]
    lib.oracle_contact_force_sum.argtypes = [vp, ci, ci, ci, ci, vp]
    lib.oracle_closest_points.argtypes = [vp, ci, ci, C.c_double, ci, vp, vp]
    lib.oracle_num_contacts.argtypes = [vp, vp, vp]
    lib.oracle_state_size.restype = C.c_size_t
    lib.oracle_state_size.argtypes = [vp]
    lib.oracle_state_get.argtypes = [vp, vp]
    lib.oracle_state_set.argtypes = [vp, vp]
    from assistive_gym_b200.capi import AgClothDesc
    lib.oracle_cloth_init.argtypes = [vp, C.POINTER(AgClothDesc)]
    lib.oracle_cloth_set_state.argtypes = [vp, vp, vp, vp]
    lib.oracle_cloth_get_state.argtypes = [vp, vp, vp]
    lib.oracle_cloth_set_anchor.argtypes = [vp, vp, vp]
    lib.oracle_cloth_anchor_follow.argtypes = [vp, ci]
    lib.oracle_cloth_set_gravity.argtypes = [vp, vp]
    lib.oracle_cloth_get_contacts.argtypes = [vp, ci, vp, vp, vp, vp, vp]
    lib.oracle_mass_matrix_inv.argtypes = [vp, ci, ci, vp]
    lib.oracle_gjk.argtypes = [vp, ci, vp, ci, vp, vp, vp]
    _libs[key] = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = np.ascontiguousarray(np.broadcast_to(a, shape))
    return a


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


class OracleSim:
    """Same method surface as assistive_gym_b200.sim.BatchSim, on the CPU oracle (double I/O)."""

    def __init__(self, scene, cfg=None, n_envs=1, f32=False, threads=1):
        from assistive_gym_b200.capi import default_config, CONTACT_DTYPE
        self.lib = _load(f32)
        self.scene = scene
        self.cfg = cfg or default_config()
        self.n = n_envs
        self.threads = threads
        self._desc = scene.as_ctypes()
        self.h = self.lib.oracle_create(C.byref(self._desc), C.byref(self.cfg), n_envs)
        if not self.h:
            raise RuntimeError('oracle_create failed')
        self.contact_dtype = CONTACT_DTYPE
        # initial base poses from the template
        for b in range(scene.n_bodies):
            self.set_base_pose(b, scene['base_pos0'][b], scene['base_quat0'][b])
        self.forward_kinematics()

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_base_pose(self, body, pos=None, quat=None, mask=None):
        pos, quat, mask = _f64(pos, (self.n, 3)), _f64(quat, (self.n, 4)), _i32(mask)
        self.lib.oracle_set_base_pose(self.h, body, _p(pos), _p(quat), _p(mask))

    def set_base_velocity(self, body, lin=None, ang=None, mask=None):
        lin, ang, mask = _f64(lin, (self.n, 3)), _f64(ang, (self.n, 3)), _i32(mask)
        self.lib.oracle_set_base_velocity(self.h, body, _p(lin), _p(ang), _p(mask))

    def set_joint_state(self, links, q=None, qd=None, mask=None):
        links = _i32(links)
        q, qd = _f64(q, (self.n, len(links))), _f64(qd, (self.n, len(links)))
        self.lib.oracle_set_joint_state(self.h, len(links), _p(links), _p(q), _p(qd), _p(_i32(mask)))

    def set_link_friction(self, link, mu, mask=None):
        mu = _f64(mu, (self.n,))
        self.lib.oracle_set_link_friction(self.h, link, _p(mu), _p(_i32(mask)))

    def set_body_active(self, body, active):
        m = _i32(np.broadcast_to(np.asarray(active, dtype=np.int32), (self.n,)))
        self.lib.oracle_set_body_mode(self.h, body, _p(m))

    def set_motor(self, links, mode, target=None, kp=None, kd=None, max_force=None):
        links = _i32(links)
        n = len(links)
        target = _f64(target, (self.n, n))
        kp = _f64(kp, (n,)) if kp is not None else None
        kd = _f64(kd if kd is not None else 1.0, (n,))
        mf = _f64(max_force, (n,)) if max_force is not None else None
        self.lib.oracle_set_motor(self.h, n, _p(links), int(mode), _p(target), _p(kp), _p(kd), _p(mf))

    def set_hard_limits(self, links, on=True):
        links = _i32(links)
        self.lib.oracle_set_hard_limits(self.h, len(links), _p(links), int(bool(on)))

    def set_motor_force_scale(self, links, scale):
        links = _i32(links)
        self.lib.oracle_set_motor_force_scale(self.h, len(links), _p(links), _p(_f64(scale, (self.n, len(links)))))

    def set_motor_targets(self, links, target):
        links = _i32(links)
        target = _f64(target, (self.n, len(links)))
        self.lib.oracle_set_motor_targets(self.h, len(links), _p(links), _p(target))

    def set_body_gravity(self, body, g):
        self.lib.oracle_set_body_gravity(self.h, int(body), _p(np.asarray(g, dtype=np.float64)))

    def forward_kinematics(self):
        self.lib.oracle_forward_kinematics(self.h)

    def step(self, n_steps=1):
        self.lib.oracle_step(self.h, n_steps, self.threads)

    def get_joint_states(self, links):
        links = _i32(links)
        n = len(links)
        q, qd, tau = (np.zeros((self.n, n)) for _ in range(3))
        self.lib.oracle_get_joint_states(self.h, n, _p(links), _p(q), _p(qd), _p(tau))
        return q, qd, tau

    def get_link_states(self, links):
        links = _i32(links)
        n = len(links)
        pos, cpos, lv, av = (np.zeros((self.n, n, 3)) for _ in range(4))
        quat, cquat = (np.zeros((self.n, n, 4)) for _ in range(2))
        self.lib.oracle_get_link_states(self.h, n, _p(links), _p(pos), _p(quat), _p(cpos), _p(cquat), _p(lv), _p(av))
        return dict(pos=pos, quat=quat, com_pos=cpos, com_quat=cquat, lin_vel=lv, ang_vel=av)

    def get_contacts(self, body_a, body_b=-2, link_a=-2, link_b=-2, max_pts=64):
        out = np.zeros((self.n, max_pts), dtype=self.contact_dtype)
        cnt = np.zeros(self.n, dtype=np.int32)
        self.lib.oracle_get_contacts(self.h, body_a, body_b, link_a, link_b, max_pts, _p(out), _p(cnt))
        return out, cnt

    def contact_force_sum(self, body_a, body_b=-2, link_a=-2, link_b=-2):
        out = np.zeros(self.n)
        self.lib.oracle_contact_force_sum(self.h, body_a, body_b, link_a, link_b, _p(out))
        return out

    def closest_points(self, body_a, body_b, distance, max_pts=64):
        out = np.zeros((self.n, max_pts), dtype=self.contact_dtype)
        cnt = np.zeros(self.n, dtype=np.int32)
        self.lib.oracle_closest_points(self.h, body_a, body_b, float(distance), max_pts, _p(out), _p(cnt))
        return out, cnt

    # ---- cloth (node arrays in the PUBLIC node order of the ClothModel)
    def cloth_init(self, model, col_links, col_static, anchor_nodes, anchor_local, gravity=(0, 0, -9.81), max_contacts=1024):
        from assistive_gym_b200 import capi
        self.cloth_model = model
        self._cloth_desc = capi.make_cloth_desc(model, self.scene, col_links, col_static, anchor_nodes, anchor_local, gravity, max_contacts)
        self.lib.oracle_cloth_init(self.h, C.byref(self._cloth_desc))

    def cloth_set_state(self, x=None, v=None, mask=None):
        m = self.cloth_model
        xi = None if x is None else _f64(m.to_internal(np.asarray(x)), (self.n, m.n_nodes, 3))
        vi = None if v is None else _f64(m.to_internal(np.asarray(v)), (self.n, m.n_nodes, 3))
        mk = None if mask is None else _i32(mask)
        self.lib.oracle_cloth_set_state(self.h, _p(xi), _p(vi), _p(mk))

    def cloth_get_state(self):
        m = self.cloth_model
        x = np.empty((self.n, m.n_nodes, 3), dtype=np.float64)
        v = np.empty_like(x)
        self.lib.oracle_cloth_get_state(self.h, _p(x), _p(v))
        return m.to_public(x), m.to_public(v)

    def cloth_set_anchor(self, pos, mask=None):
        mk = None if mask is None else _i32(mask)
        self.lib.oracle_cloth_set_anchor(self.h, _p(_f64(pos, (self.n, 3))), _p(mk))

    def cloth_anchor_follow(self, link):
        self.lib.oracle_cloth_anchor_follow(self.h, int(link))

    def cloth_set_gravity(self, g):
        self.lib.oracle_cloth_set_gravity(self.h, _p(np.asarray(g, dtype=np.float64)))

    def cloth_get_contacts(self, max_pts=1024):
        cnt = np.zeros(self.n, dtype=np.int32)
        node = np.zeros((self.n, max_pts), dtype=np.int32)
        link = np.zeros((self.n, max_pts), dtype=np.int32)
        pos = np.zeros((self.n, max_pts, 3), dtype=np.float64)
        force = np.zeros((self.n, max_pts, 3), dtype=np.float64)
        self.lib.oracle_cloth_get_contacts(self.h, max_pts, _p(cnt), _p(node), _p(pos), _p(force), _p(link))
        return cnt, self.cloth_model.order[node], pos, force, link

    def num_contacts(self):
        cnt, it = np.zeros(self.n, dtype=np.int32), np.zeros(self.n, dtype=np.int32)
        self.lib.oracle_num_contacts(self.h, _p(cnt), _p(it))
        return cnt, it

    def state_get(self):
        sz = self.lib.oracle_state_size(self.h)
        out = np.zeros((self.n, sz))
        self.lib.oracle_state_get(self.h, _p(out))
        return out

    def state_set(self, st):
        st = _f64(st)
        self.lib.oracle_state_set(self.h, _p(st))

    def mass_matrix_inv(self, body, env=0):
        nd = self.lib.oracle_num_dofs(self.h, body)
        out = np.zeros((nd, nd))
        self.lib.oracle_mass_matrix_inv(self.h, env, body, _p(out))
        return out


def gjk(A, B, f32=False):
    lib = _load(f32)
    A, B = _f64(A), _f64(B)
    pa, pb, d = np.zeros(3), np.zeros(3), np.zeros(1)
    ov = lib.oracle_gjk(_p(A), len(A), _p(B), len(B), _p(pa), _p(pb), _p(d))
    return bool(ov), pa, pb, float(d[0])
