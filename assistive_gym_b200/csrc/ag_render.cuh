// ag_render.cuh — K9: camera images of the batched scenes (reference envs/env.py:342-359: setup_camera / setup_camera_rpy ->
// p.computeViewMatrix, p.computeProjectionMatrixFOV; get_camera_image_depth -> p.getCameraImage; used by learn.py:101,125 to
// write the --colab APNG).  PyBullet rasterises the VISUAL meshes with OpenGL / TinyRenderer; the visual meshes are not part
// of the compiled scene (SURVEY.md section 2 keeps the viewer out of scope), so this kernel ray-casts the COLLISION geometry:
// one thread per (pixel, env), analytic ray-sphere / ray-capsule intersections, ray-polytope clipping against the face planes
// of hull colliders (pushed out by the collision margin), Lambert shading with PyBullet's ambient / diffuse coefficients and a
// per-body colour; the depth image is the OpenGL depth-buffer value for the same near / far planes, as getCameraImage returns.
#pragma once
#include "ag_device.cuh"
#include "../../include/agphys.h"

struct RenderDev {
  AgCamera cam;
  f3 fwd, right, up;            // camera basis (computeViewMatrix)
  float tan_half;               // tan(fov / 2)
  const int* env_ids;           // [n_render]
  unsigned char* rgba;          // [n_render][H][W][4]
  float* depth;                 // [n_render][H][W]
};

// nearest hit of the ray o + t d (link frame, |d| = 1) with collider c for t in (t0, t1); returns t (or t1) and the normal
AG_HDN inline float render_hit(const SimDev& S, int c, f3 o, f3 d, float t0, float t1, f3& nrm) {
  int type = AG_LDG(S.col_type + c), v0 = AG_LDG(S.col_v0 + c);
  float r = AG_LDG(S.col_radius + c);
  if (type == 0 || type == 1) {
    f3 a = tv3(S.verts, v0);
    float best = t1;
    // sphere caps (a capsule = two spheres + a cylinder between them)
    for (int k = 0; k <= type; k++) {
      f3 cc = k == 0 ? a : tv3(S.verts, v0 + 1);
      f3 m = o - cc;
      float b = dot(m, d), cq = dot(m, m) - r * r, disc = b * b - cq;
      if (disc >= 0.f) { float t = -b - sqrtf(disc); if (t > t0 && t < best) { best = t; f3 p = o + d * t; nrm = (p - cc) * (1.f / r); } }
    }
    if (type == 1) {
      f3 b1 = tv3(S.verts, v0 + 1), ax = b1 - a;
      float L2 = dot(ax, ax);
      if (L2 > 1e-16f) {
        float L = sqrtf(L2); f3 u = ax * (1.f / L);
        f3 m = o - a;
        f3 dp = d - u * dot(d, u), mp = m - u * dot(m, u);
        float A = dot(dp, dp), B = dot(dp, mp), Cq = dot(mp, mp) - r * r, disc = B * B - A * Cq;
        if (A > 1e-12f && disc >= 0.f) {
          float t = (-B - sqrtf(disc)) / A;
          float h = dot(m + d * t, u);
          if (t > t0 && t < best && h >= 0.f && h <= L) { best = t; f3 p = m + d * t; nrm = (p - u * h) * (1.f / r); }
        }
      }
    }
    return best;
  }
  // hull / half-space: clip the ray against every face plane n.x <= dpl + r
  int p0 = AG_LDG(S.col_p0 + c), np = AG_LDG(S.col_np + c);
  float te = t0, tx = t1; f3 ne(0.f, 0.f, 1.f); bool entered = false;
  for (int k = p0; k < p0 + np; k++) {
    f3 pn; float pd; ld_plane(S.planes, k, pn, pd);
    float denom = dot(pn, d), dist = dot(pn, o) - (pd + r);
    if (fabsf(denom) < 1e-9f) { if (dist > 0.f) return t1; continue; }
    float t = -dist / denom;
    if (denom < 0.f) { if (t > te) { te = t; ne = pn; entered = true; } } else if (t < tx) tx = t;
    if (te > tx) return t1;
  }
  if (!entered || np == 0) return t1;
  nrm = ne;
  return te;
}

// thread = (pixel, env slot), env slot fastest is NOT used here: pixels of one env are contiguous (coalesced image writes)
AG_HDN inline void render_body(int tid, const SimDev& S, const KP& p) {
  const RenderDev& R = *(const RenderDev*)p.p0;
  const int W = R.cam.width, H = R.cam.height, N = S.N;
  const int pix = tid % (W * H), slot = tid / (W * H);
  const int e = R.env_ids[slot];
  const int row = pix / W, col = pix % W;
  float xn = (2.f * (col + 0.5f) / W - 1.f) * R.tan_half * R.cam.aspect, yn = (1.f - 2.f * (row + 0.5f) / H) * R.tan_half;
  f3 dirv = R.fwd + R.right * xn + R.up * yn;
  float dl = norm(dirv);
  f3 d = dirv * (1.f / dl);
  f3 o(R.cam.eye[0], R.cam.eye[1], R.cam.eye[2]);
  const float cosf_ = 1.f / dl;                       // cos of the angle to the optical axis: z_eye = t * cos
  float tbest = R.cam.far_ / cosf_; f3 nbest(0.f, 0.f, 1.f); int cbest = -1;
  const float tmin = R.cam.near_ / cosf_;
  for (int k = 0; k < S.nl; k++) {
    int nc = AG_LDG(S.link_ncol + k);
    if (nc == 0) continue;
    if (S.body_mode[(size_t)AG_LDG(S.link_body + k) * N + e] == 0) continue;
    // bounding box of the link (world AABB kept by the step) against the ray: slab test
    f3 lo = ld3(S.lmin, k, N, e), hi = ld3(S.lmax, k, N, e);
    float t0 = tmin, t1 = tbest; bool miss = false;
    for (int a = 0; a < 3; a++) {
      float oa = comp(o, a), da = comp(d, a), l = comp(lo, a), h = comp(hi, a);
      if (fabsf(da) < 1e-12f) { if (oa < l || oa > h) miss = true; continue; }
      float ta = (l - oa) / da, tb = (h - oa) / da;
      if (ta > tb) { float sw = ta; ta = tb; tb = sw; }
      t0 = fmaxf(t0, ta); t1 = fminf(t1, tb);
    }
    if (miss || t0 > t1) continue;
    f3 lp = ld3(S.lpos, k, N, e); q4 lq = ld4(S.lquat, k, N, e);
    f3 ol = qrot_inv(lq, o - lp), dloc = qrot_inv(lq, d);
    int c0 = AG_LDG(S.link_col0 + k);
    for (int c = c0; c < c0 + nc; c++) {
      f3 n;
      float t = render_hit(S, c, ol, dloc, tmin, tbest, n);
      if (t < tbest) { tbest = t; nbest = qrot(lq, n); cbest = c; }
    }
  }
  unsigned char* px = R.rgba + ((size_t)slot * H * W + pix) * 4;
  float zb = 1.f;
  if (cbest >= 0) {
    float ze = tbest * cosf_;
    float nr = R.cam.near_, fr = R.cam.far_;
    zb = 0.5f * ((fr + nr) / (fr - nr) - 2.f * fr * nr / ((fr - nr) * ze)) + 0.5f;
    // a fixed palette by body: shades of the reference's default materials are not part of the scene description
    int b = AG_LDG(S.link_body + AG_LDG(S.col_link + cbest));
    const float pal[8][3] = {{0.85f, 0.85f, 0.85f}, {0.55f, 0.65f, 0.85f}, {0.95f, 0.75f, 0.6f}, {0.95f, 0.75f, 0.6f},
                             {0.35f, 0.35f, 0.4f}, {0.8f, 0.6f, 0.3f}, {0.7f, 0.7f, 0.75f}, {0.9f, 0.9f, 0.5f}};
    f3 ld(R.cam.light_dir[0], R.cam.light_dir[1], R.cam.light_dir[2]);
    float ll = norm(ld); if (ll > 0.f) ld = ld * (1.f / ll);
    float lam = fmaxf(dot(nbest, ld), 0.f);
    float sh = fminf(1.f, R.cam.ambient + R.cam.diffuse * lam);
    for (int a = 0; a < 3; a++) px[a] = (unsigned char)(255.f * fminf(1.f, pal[b & 7][a] * sh) + 0.5f);
    px[3] = 255;
  } else { px[0] = px[1] = px[2] = 255; px[3] = 255; }
  R.depth[(size_t)slot * H * W + pix] = zb;
}
