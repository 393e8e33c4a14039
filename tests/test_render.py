"""Camera images (K9; reference envs/env.py:342-359): the device ray caster against a numpy
restatement that works in WORLD space on the oracle's link poses (the kernel transforms each ray into the link frame)."""
import numpy as np
import pytest

from assistive_gym_b200 import capi
from assistive_gym_b200.feeding_batch import FeedingBatch
from assistive_gym_b200.kinematics import q_rot
from assistive_gym_b200.sim import BatchSim
from oracle.oracle_py import OracleSim

W, H = 96, 54
EYE, TARGET, FOV, NEAR, FAR = np.array([0.5, -0.75, 1.5]), np.array([-0.2, 0, 0.75]), 60.0, 0.01, 100.0       # env.py:342


def _rays():
    f = TARGET - EYE; f = f / np.linalg.norm(f)
    r = np.cross(f, [0, 0, 1.0]); r = r / np.linalg.norm(r)
    u = np.cross(r, f)
    th = np.tan(np.deg2rad(FOV) / 2)
    col, row = np.meshgrid(np.arange(W), np.arange(H))
    xn = (2 * (col + 0.5) / W - 1) * th * (W / H)
    yn = (1 - 2 * (row + 0.5) / H) * th
    d = f[None, None] + xn[..., None] * r + yn[..., None] * u
    dl = np.linalg.norm(d, axis=-1, keepdims=True)
    return d / dl, 1.0 / dl[..., 0]


def reference_depth(scene, sim, env=0, skip_bodies=()):
    """World-space ray casting of every collider of env `env`: returns the depth-buffer image and the id of the hit collider."""
    sc = scene
    d, cosv = _rays()
    tbest = np.full((H, W), FAR) / cosv
    hit = np.full((H, W), -1)
    ls = sim.get_link_states(list(range(sc.n_links)))
    lp, lq = ls['pos'][env].astype(np.float64), ls['quat'][env].astype(np.float64)
    for c in range(sc.n_colliders):
        k = int(sc['col_link'][c])
        if int(sc['link_body'][k]) in skip_bodies:          # a body switched off in this env (the other-gender person)
            continue
        v = sc['verts'][sc['col_v0'][c]:sc['col_v0'][c] + sc['col_nv'][c]]
        vw = q_rot(lq[k][None], v) + lp[k]
        r = float(sc['col_radius'][c])
        t = np.full((H, W), np.inf)
        if sc['col_type'][c] in (0, 1):
            for cc in vw:
                m = EYE - cc
                b = d @ m; cq = m @ m - r * r; disc = b * b - cq
                tt = np.where(disc >= 0, -b - np.sqrt(np.maximum(disc, 0)), np.inf)
                t = np.minimum(t, np.where(tt > NEAR / cosv, tt, np.inf))
            if sc['col_type'][c] == 1:
                a, b1 = vw
                ax = b1 - a; L = np.linalg.norm(ax); uax = ax / L
                m = EYE - a
                dp = d - (d @ uax)[..., None] * uax; mp = m - (m @ uax) * uax
                A = (dp * dp).sum(-1); B = dp @ mp; Cq = mp @ mp - r * r; disc = B * B - A * Cq
                with np.errstate(divide='ignore', invalid='ignore'):
                    tt = np.where((disc >= 0) & (A > 1e-12), (-B - np.sqrt(np.maximum(disc, 0))) / A, np.inf)
                h = (m[None, None] + d * tt[..., None]) @ uax
                ok = (tt > NEAR / cosv) & (h >= 0) & (h <= L)
                t = np.minimum(t, np.where(ok, tt, np.inf))
        else:
            pl = sc['planes'][sc['col_p0'][c]:sc['col_p0'][c] + sc['col_np'][c]]
            if len(pl) == 0:
                continue
            n = q_rot(lq[k][None], pl[:, :3]); dpl = pl[:, 3] + n @ lp[k] + r
            te = np.full((H, W), -np.inf); tx = np.full((H, W), np.inf); par_out = np.zeros((H, W), bool)
            for nn, dd in zip(n, dpl):
                denom = d @ nn; dist = EYE @ nn - dd
                with np.errstate(divide='ignore', invalid='ignore'):
                    tt = -dist / denom
                par_out |= (np.abs(denom) < 1e-9) & (dist > 0)
                te = np.where(denom < -1e-9, np.maximum(te, tt), te)
                tx = np.where(denom > 1e-9, np.minimum(tx, tt), tx)
            ok = (te <= tx) & ~par_out & np.isfinite(te) & (te > NEAR / cosv)
            t = np.where(ok, te, np.inf)
        better = t < tbest
        tbest = np.where(better, t, tbest); hit = np.where(better, c, hit)
    ze = tbest * cosv
    zb = np.where(hit >= 0, 0.5 * ((FAR + NEAR) / (FAR - NEAR) - 2 * FAR * NEAR / ((FAR - NEAR) * ze)) + 0.5, 1.0)
    return zb, hit


def _check(lib):
    fb = FeedingBatch()
    sim = BatchSim(fb.scene, capi.default_config(), 2, _lib=lib)
    smp = fb.reset(sim, np.random.default_rng(4), settle_steps=0)
    img, depth = sim.render(EYE, TARGET, fov=FOV, width=W, height=H, env_ids=[0, 1])
    assert img.shape == (2, H, W, 4) and img.dtype == np.uint8 and depth.shape == (2, H, W)
    for e in range(2):
        zb, hit = reference_depth(fb.scene, sim, env=e, skip_bodies=(fb.humans['female' if smp['male'][e] else 'male'],))
        bg_ref, bg = hit < 0, depth[e] >= 1.0
        assert (bg_ref != bg).mean() < 0.01                                   # silhouettes may differ by a pixel
        both = ~bg_ref & ~bg
        err = np.abs(depth[e] - zb)[both]
        assert both.mean() > 0.5 and np.median(err) < 1e-6 and (err > 1e-4).mean() < 0.02, (both.mean(), np.median(err), (err > 1e-4).mean())
        assert np.all(img[e][bg] == 255) and img[e][both][:, :3].min() < 250 and np.all(img[e][..., 3] == 255)
    assert not np.array_equal(depth[0], depth[1])                             # the two envs differ (gender, head pose, bowl, arm)
    sim.close()


def test_render_host_compiled_kernel_body(emu_lib):
    _check(emu_lib)


@pytest.mark.gpu
def test_render_cuda(gpu_lib):
    _check(gpu_lib)


def test_env_camera_api(emu_lib):
    """setup_camera / get_camera_image_depth / render('rgb_array') with the reference's signatures (env.py:342-359, learn.py:101-125)."""
    from assistive_gym_b200 import envs
    env = envs.make('FeedingJaco-v1', n_envs=1)
    env._sim_lib = emu_lib
    env.reset()
    env.setup_camera(camera_eye=[0.5, -0.75, 1.5], camera_target=[-0.2, 0, 0.75], fov=60, camera_width=64, camera_height=36)
    img, depth = env.get_camera_image_depth()
    assert img.shape == (36, 64, 4) and depth.shape == (36, 64) and depth.min() < 1.0
    env.setup_camera_rpy(camera_width=64, camera_height=36)
    img2 = env.render(mode='rgb_array')
    assert img2.shape == (36, 64, 4) and not np.array_equal(img, img2)
    env.close()
