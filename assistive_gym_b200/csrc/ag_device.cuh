// ag_device.cuh — per-lane bodies of the sm_100a kernels (one env per lane, lock-step).
//
// Every function here is `__host__ __device__`: the CUDA build wraps them in __global__ kernels
// (agphys.cu); tests/kernel_harness compiles the same bodies for the host to check kernel logic
// against the CPU oracle on a box without a GPU.  The harness is a test aid — the package only
// ever loads the CUDA library.
//
// Reference call sites this path replaces: p.stepSimulation (envs/env.py:226, feeding.py:179) and the
// read-back calls in envs/agents/agent.py:40,52,72,108,124.  The algorithms restate what Bullet does
// for that call: collision detection at the current pose, Featherstone ABA
// for the unconstrained velocity update, velocity-level PGS over joint limits, joint motors, fixed
// constraints and frictional contacts, symplectic Euler.
#pragma once
#include <string.h>
#include "ag_math.cuh"
#include "ag_types.h"

// ------------------------------------------------------------------ SoA accessors
AG_HD float ld1(const float* p, int item, int N, int e) { return p[(size_t)item * N + e]; }
AG_HD void st1(float* p, int item, int N, int e, float v) { p[(size_t)item * N + e] = v; }
AG_HD f3 ld3(const float* p, int item, int N, int e) {
  size_t b = (size_t)item * 3 * N + e;
  return f3(p[b], p[b + N], p[b + 2 * (size_t)N]);
}
AG_HD void st3(float* p, int item, int N, int e, f3 v) {
  size_t b = (size_t)item * 3 * N + e;
  p[b] = v.x; p[b + N] = v.y; p[b + 2 * (size_t)N] = v.z;
}
AG_HD q4 ld4(const float* p, int item, int N, int e) {
  size_t b = (size_t)item * 4 * N + e;
  return q4(p[b], p[b + N], p[b + 2 * (size_t)N], p[b + 3 * (size_t)N]);
}
AG_HD void st4(float* p, int item, int N, int e, q4 v) {
  size_t b = (size_t)item * 4 * N + e;
  p[b] = v.x; p[b + N] = v.y; p[b + 2 * (size_t)N] = v.z; p[b + 3 * (size_t)N] = v.w;
}
AG_HD f3 tv3(const float* p, int i) { return f3(AG_LDG(p + 3 * i), AG_LDG(p + 3 * i + 1), AG_LDG(p + 3 * i + 2)); }
AG_HD q4 tv4(const float* p, int i) { return q4(AG_LDG(p + 4 * i), AG_LDG(p + 4 * i + 1), AG_LDG(p + 4 * i + 2), AG_LDG(p + 4 * i + 3)); }
AG_HD float cf_ld(const float* d, int slot, int f, int N, int e) { return d[((size_t)slot * AG_CF + f) * N + e]; }
AG_HD void cf_st(float* d, int slot, int f, int N, int e, float v) { d[((size_t)slot * AG_CF + f) * N + e] = v; }

// plane k of collider (local) -> (n, d)
AG_HD void ld_plane(const float* planes, int k, f3& n, float& d) {
  n = f3(AG_LDG(planes + 4 * k), AG_LDG(planes + 4 * k + 1), AG_LDG(planes + 4 * k + 2)); d = AG_LDG(planes + 4 * k + 3);
}

// ------------------------------------------------------------------ K1: forward kinematics
// One lane per env.  p.i0 != 0: all bodies (reset / after teleports); else only movable bodies.
AG_HDN inline void fk_body(int e, const SimDev& S, const KP& p) {
  const int N = S.N;
  for (int b = 0; b < S.nb; b++) {
    if (!p.i0 && AG_LDG(S.body_kind + b) == BK_STATIC) continue;
    int l0 = AG_LDG(S.body_link0 + b), nlk = AG_LDG(S.body_nlinks + b);
    f3 bp = ld3(S.base_pos, b, N, e);
    q4 bq = ld4(S.base_quat, b, N, e);
    st3(S.lpos, l0, N, e, bp); st4(S.lquat, l0, N, e, bq);
    for (int k = l0 + 1; k < l0 + nlk; k++) {
      int par = AG_LDG(S.link_parent + k);
      f3 pp = ld3(S.lpos, par, N, e);
      q4 pq = ld4(S.lquat, par, N, e);
      f3 jp = pp + qrot(pq, tv3(S.link_jpos, k));
      q4 jq = qmul(pq, tv4(S.link_jquat, k));
      int jt = AG_LDG(S.link_jtype + k);
      if (jt == 1) jq = qmul(jq, qaxis(tv3(S.link_axis, k), ld1(S.jq, k, N, e)));
      else if (jt == 2) jp = jp + qrot(jq, tv3(S.link_axis, k) * ld1(S.jq, k, N, e));
      st3(S.lpos, k, N, e, jp); st4(S.lquat, k, N, e, qnormalize(jq));
    }
  }
}

// ------------------------------------------------------------------ K2a: collider AABBs
// thread = (collider list index i, env e), env fastest.  p.p0 = list, p.i0 = list length.
AG_HDN inline void aabb_body(int tid, const SimDev& S, const KP& p) {
  const int N = S.N;
  int e = tid % N, i = tid / N;
  int c = AG_LDG((const int*)p.p0 + i);
  int k = AG_LDG(S.col_link + c);
  if (AG_LDG(S.col_type + c) == 3) {   // half-space: unbounded, except along an axis-aligned normal
    f3 n; float d; ld_plane(S.planes, AG_LDG(S.col_p0 + c), n, d);
    q4 q = ld4(S.lquat, k, N, e);
    f3 nw = qrot(q, n);
    float dw = d + dot(nw, ld3(S.lpos, k, N, e));
    f3 mn(-1e30f, -1e30f, -1e30f), mx(1e30f, 1e30f, 1e30f);
    if (nw.x > 0.999999f) mx.x = dw; else if (nw.x < -0.999999f) mn.x = -dw;
    if (nw.y > 0.999999f) mx.y = dw; else if (nw.y < -0.999999f) mn.y = -dw;
    if (nw.z > 0.999999f) mx.z = dw; else if (nw.z < -0.999999f) mn.z = -dw;
    st3(S.cmin, c, N, e, mn); st3(S.cmax, c, N, e, mx);
    return;
  }
  f3 lp = ld3(S.lpos, k, N, e);
  m3 R = qmat(ld4(S.lquat, k, N, e));
  f3 ctr = lp + mul(R, tv3(S.col_center, c));
  f3 h = tv3(S.col_half, c);
  float r = AG_LDG(S.col_radius + c);
  f3 hw(fabsf(R.m[0]) * h.x + fabsf(R.m[1]) * h.y + fabsf(R.m[2]) * h.z + r,
        fabsf(R.m[3]) * h.x + fabsf(R.m[4]) * h.y + fabsf(R.m[5]) * h.z + r,
        fabsf(R.m[6]) * h.x + fabsf(R.m[7]) * h.y + fabsf(R.m[8]) * h.z + r);
  st3(S.cmin, c, N, e, ctr - hw); st3(S.cmax, c, N, e, ctr + hw);
}
// K2b: link AABBs = union over the link's colliders.  thread = (link list index, env).
AG_HDN inline void linkaabb_body(int tid, const SimDev& S, const KP& p) {
  const int N = S.N;
  int e = tid % N, i = tid / N;
  int k = AG_LDG((const int*)p.p0 + i);
  int c0 = AG_LDG(S.link_col0 + k), ncl = AG_LDG(S.link_ncol + k);
  f3 mn(1e30f, 1e30f, 1e30f), mx(-1e30f, -1e30f, -1e30f);
  for (int c = c0; c < c0 + ncl; c++) { mn = fmin3(mn, ld3(S.cmin, c, N, e)); mx = fmax3(mx, ld3(S.cmax, c, N, e)); }
  st3(S.lmin, k, N, e, mn); st3(S.lmax, k, N, e, mx);
}

// ------------------------------------------------------------------ K3: narrowphase
AG_HD bool aabb_ov(f3 amin, f3 amax, f3 bmin, f3 bmax, float m) {
  return !(amin.x > bmax.x + m || bmin.x > amax.x + m || amin.y > bmax.y + m || bmin.y > amax.y + m ||
           amin.z > bmax.z + m || bmin.z > amax.z + m);
}

// closest point on segment / triangle to the origin (barycentric), Ericson RTCD 5.1 — in fp64:
// the simplex vertices are differences of support points that can be ~1 m apart while the origin is
// ~1 mm from the simplex; in fp32 the Voronoi-region determinants lose all significance on such thin
// simplices (measured on B200: 9 % of link-vs-table-edge queries off by up to 2 mm).
AG_HD void seg_origin(d3 a, d3 b, double& u, double& v) {
  d3 ab = b - a;
  double t = -dot(a, ab), den = dot(ab, ab);
  if (t <= 0.0 || den <= 0.0) { u = 1.0; v = 0.0; return; }
  if (t >= den) { u = 0.0; v = 1.0; return; }
  v = t / den; u = 1.0 - v;
}
AG_HD void tri_origin(d3 a, d3 b, d3 c, double& u, double& v, double& w) {
  d3 ab = b - a, ac = c - a;
  double d1 = -dot(ab, a), d2 = -dot(ac, a);
  if (d1 <= 0.0 && d2 <= 0.0) { u = 1.0; v = 0.0; w = 0.0; return; }
  double d3_ = -dot(ab, b), d4 = -dot(ac, b);
  if (d3_ >= 0.0 && d4 <= d3_) { u = 0.0; v = 1.0; w = 0.0; return; }
  double vc = d1 * d4 - d3_ * d2;
  if (vc <= 0.0 && d1 >= 0.0 && d3_ <= 0.0) { double t = d1 / (d1 - d3_); u = 1.0 - t; v = t; w = 0.0; return; }
  double d5 = -dot(ab, c), d6 = -dot(ac, c);
  if (d6 >= 0.0 && d5 <= d6) { u = 0.0; v = 0.0; w = 1.0; return; }
  double vb = d5 * d2 - d1 * d6;
  if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) { double t = d2 / (d2 - d6); u = 1.0 - t; v = 0.0; w = t; return; }
  double va = d3_ * d6 - d5 * d4;
  if (va <= 0.0 && (d4 - d3_) >= 0.0 && (d5 - d6) >= 0.0) { double t = (d4 - d3_) / ((d4 - d3_) + (d5 - d6)); u = 0.0; v = 1.0 - t; w = t; return; }
  double den = 1.0 / (va + vb + vc);
  v = vb * den; w = vc * den; u = 1.0 - v - w;
}

struct NpOut { f3 pa, pb, n; float d; };   // B-local frame: points on the surfaces, normal B->A, surface distance

// GJK closest points between core A (A-local vertices mapped by R,t into B's frame) and core B.
// Vertices, transforms and the support search are fp32; the simplex solve is fp64.
// Returns true if the cores overlap.
AG_HDN inline bool gjk_cores(const float* verts, int va0, int nA, int vb0, int nB, const m3& R, f3 t,
                             f3& pa, f3& pb, f3& nrm, float& dist) {
  d3 W[4]; f3 PA[4], PB[4];
  int IA[4], IB[4];
  double lam[4] = {1.0, 0.0, 0.0, 0.0};
  int n = 0;
  d3 v = to_d3(mul(R, tv3(verts, va0)) + t) - to_d3(tv3(verts, vb0));
  if (dot(v, v) < 1e-20) v = d3(1.0, 0.0, 0.0);
  bool overlap = false;
  double lower_bound = 0.0;
  for (int it = 0; it < 32; it++) {
    // support of A in direction -v (A-local: R^T(-v)), support of B in +v
    f3 vf = to_f3(v);
    f3 da = mulT(R, -vf);
    int ia = 0, ib = 0;
    float best = dot(da, tv3(verts, va0));
    for (int i = 1; i < nA; i++) { float d = dot(da, tv3(verts, va0 + i)); if (d > best) { best = d; ia = i; } }
    best = dot(vf, tv3(verts, vb0));
    for (int i = 1; i < nB; i++) { float d = dot(vf, tv3(verts, vb0 + i)); if (d > best) { best = d; ib = i; } }
    f3 sa = mul(R, tv3(verts, va0 + ia)) + t;
    f3 sb = tv3(verts, vb0 + ib);
    d3 w = to_d3(sa) - to_d3(sb);
    double vv = dot(v, v);
    double vw = dot(v, w);
    if (n > 0 && vw > 0.0) lower_bound = fmax(lower_bound, vw / sqrt(vv));
    if (n > 0 && vv - vw <= 1e-10 * vv) break;
    bool dup = false;
    for (int i = 0; i < n; i++) if (IA[i] == ia && IB[i] == ib) dup = true;
    if (dup) break;
    W[n] = w; PA[n] = sa; PB[n] = sb; IA[n] = ia; IB[n] = ib; n++;
    if (n == 1) { lam[0] = 1.0; }
    else if (n == 2) {
      double u, s; seg_origin(W[0], W[1], u, s);
      if (s <= 0.0) { n = 1; lam[0] = 1.0; }
      else if (u <= 0.0) { W[0] = W[1]; PA[0] = PA[1]; PB[0] = PB[1]; IA[0] = IA[1]; IB[0] = IB[1]; n = 1; lam[0] = 1.0; }
      else { lam[0] = u; lam[1] = s; }
    } else if (n == 3) {
      double l3[3]; tri_origin(W[0], W[1], W[2], l3[0], l3[1], l3[2]);
      int m = 0;
      for (int i = 0; i < 3; i++) if (l3[i] > 0.0) { W[m] = W[i]; PA[m] = PA[i]; PB[m] = PB[i]; IA[m] = IA[i]; IB[m] = IB[i]; lam[m] = l3[i]; m++; }
      n = m;
    } else {
      double bestd = 1e300; int bf = -1; double bl[3] = {0.0, 0.0, 0.0};
      double bestd_all = 1e300; int bf_all = 0; double bla[3] = {1.0, 0.0, 0.0};
      bool any_out = false;
      for (int f = 0; f < 4; f++) {
        int i0 = (f == 3) ? 1 : 0, i1 = (f == 0) ? 1 : ((f == 1) ? 2 : 3), i2 = (f == 0) ? 2 : ((f == 1) ? 3 : ((f == 2) ? 1 : 2)), i3 = (f == 0) ? 3 : ((f == 1) ? 1 : ((f == 2) ? 2 : 0));
        d3 a = W[i0], b = W[i1], c = W[i2], d = W[i3];
        d3 nn = cross(b - a, c - a);
        double sp = -dot(a, nn), sd = dot(d - a, nn);
        // inside only if CLEARLY on the opposite vertex's side; flat tetrahedra count as outside
        double nl = sqrt(dot(nn, nn));
        double tol_d = 1e-9 * nl * sqrt(dot(d - a, d - a)), tol_p = 1e-9 * nl * sqrt(dot(a, a));
        bool inside = (sp * sd > 0.0) && (fabs(sd) > tol_d) && (fabs(sp) > tol_p);
        double u, s, r; tri_origin(a, b, c, u, s, r);
        d3 pt = a * u + b * s + c * r;
        double dd = dot(pt, pt);
        if (dd < bestd_all) { bestd_all = dd; bf_all = f; bla[0] = u; bla[1] = s; bla[2] = r; }
        if (inside) continue;
        any_out = true;
        if (dd < bestd) { bestd = dd; bf = f; bl[0] = u; bl[1] = s; bl[2] = r; }
      }
      if (!any_out) {
        // a positive lower bound on the distance (v.w/|v| of an earlier iteration) proves separation
        if (lower_bound <= 1e-7) { overlap = true; break; }
        bf = bf_all; bl[0] = bla[0]; bl[1] = bla[1]; bl[2] = bla[2];
      }
      int f = bf;
      int id[3];
      id[0] = (f == 3) ? 1 : 0; id[1] = (f == 0) ? 1 : ((f == 1) ? 2 : 3); id[2] = (f == 0) ? 2 : ((f == 1) ? 3 : ((f == 2) ? 1 : 2));
      d3 tw[3]; f3 ta[3], tb[3]; int tia[3], tib[3];
      for (int i = 0; i < 3; i++) { tw[i] = W[id[i]]; ta[i] = PA[id[i]]; tb[i] = PB[id[i]]; tia[i] = IA[id[i]]; tib[i] = IB[id[i]]; }
      int m = 0;
      for (int i = 0; i < 3; i++) if (bl[i] > 0.0) { W[m] = tw[i]; PA[m] = ta[i]; PB[m] = tb[i]; IA[m] = tia[i]; IB[m] = tib[i]; lam[m] = bl[i]; m++; }
      n = m;
    }
    d3 nv(0.0, 0.0, 0.0);
    for (int i = 0; i < n; i++) nv = nv + W[i] * lam[i];
    v = nv;
    if (dot(v, v) <= 1e-16) { overlap = true; break; }
  }
  if (overlap) return true;
  d3 qa(0.0, 0.0, 0.0), qb(0.0, 0.0, 0.0);
  for (int i = 0; i < n; i++) { qa = qa + to_d3(PA[i]) * lam[i]; qb = qb + to_d3(PB[i]) * lam[i]; }
  d3 d = qa - qb;
  double dn = sqrt(dot(d, d));
  pa = to_f3(qa); pb = to_f3(qb);
  dist = (float)dn;
  nrm = dn > 0.0 ? to_f3(d * (1.0 / dn)) : f3(0.f, 0.f, 1.f);
  return false;
}

// axis of least penetration over the face normals of both cores (B-local frame)
AG_HDN inline void pen_faces(const SimDev& S, int va0, int nA, int pa0, int npA, int vb0, int nB, int pb0, int npB,
                             const m3& R, f3 t, f3& pa, f3& pb, f3& nrm, float& dist) {
  float best = -1e30f; bool found = false;
  for (int k = 0; k < npA; k++) {
    f3 nl; float dl; ld_plane(S.planes, pa0 + k, nl, dl);
    f3 n = mul(R, nl); float d = dl + dot(n, t);
    int jb = 0; float mn = dot(n, tv3(S.verts, vb0));
    for (int j = 1; j < nB; j++) { float x = dot(n, tv3(S.verts, vb0 + j)); if (x < mn) { mn = x; jb = j; } }
    float sep = mn - d;
    if (sep > best) { best = sep; found = true; nrm = -n; pb = tv3(S.verts, vb0 + jb); pa = pb - n * sep; }
  }
  for (int k = 0; k < npB; k++) {
    f3 n; float d; ld_plane(S.planes, pb0 + k, n, d);
    f3 nl = mulT(R, n);
    int ja = 0; float mn = dot(nl, tv3(S.verts, va0));
    for (int j = 1; j < nA; j++) { float x = dot(nl, tv3(S.verts, va0 + j)); if (x < mn) { mn = x; ja = j; } }
    float sep = mn + dot(n, t) - d;
    if (sep > best) { best = sep; found = true; nrm = n; pa = mul(R, tv3(S.verts, va0 + ja)) + t; pb = pa - n * sep; }
  }
  if (!found) { nrm = f3(0.f, 0.f, 1.f); pa = mul(R, tv3(S.verts, va0)) + t; pb = tv3(S.verts, vb0); best = 0.f; }
  dist = fminf(best, 0.f);
}

struct CandSet { NpOut c[4]; int n; };
// keep the primary + up to 3 more, greedily the farthest from the chosen set
struct CandSel {
  NpOut prim; NpOut pool[12]; int np;
};

// vertices of core V that lie over the supporting face of the plane owner (see oracle for the rule).
// Everything in B-local coordinates; `planes are given by (p0, np, xf)` where xf says whether plane
// normals need mapping by R,t (owner is A) or not (owner is B).
AG_HDN inline void face_cands(const SimDev& S, int vv0, int nV, float rV, bool v_is_a, int p0, int np, float rP,
                              const m3& R, f3 t, f3 n_to_v, float d_primary, float tol, float max_dist, CandSel& cs) {
  // supporting face
  int kf = -1; float best = 0.98f; f3 nf; float df = 0.f;
  for (int k = 0; k < np; k++) {
    f3 n; float d; ld_plane(S.planes, p0 + k, n, d);
    if (!v_is_a) { f3 nw = mul(R, n); d = d + dot(nw, t); n = nw; }   // plane owner is A: map to B-local
    float al = dot(n, n_to_v);
    if (al > best) { best = al; kf = k; nf = n; df = d; }
  }
  if (kf < 0) return;
  for (int i = 0; i < nV; i++) {
    f3 v = tv3(S.verts, vv0 + i);
    if (v_is_a) v = mul(R, v) + t;
    float h = dot(nf, v) - df;
    float d = h - rV - rP;
    if (d > d_primary + tol || d > max_dist) continue;
    f3 proj = v - nf * h;
    bool inside = true;
    for (int k = 0; k < np; k++) {
      if (k == kf) continue;
      f3 n; float dd; ld_plane(S.planes, p0 + k, n, dd);
      if (!v_is_a) { f3 nw = mul(R, n); dd = dd + dot(nw, t); n = nw; }
      if (dot(n, proj) - dd > 1e-6f) { inside = false; break; }
    }
    if (!inside) continue;
    if (cs.np >= 12) {
      // pool full: replace the shallowest entry if this one is deeper
      int wi = 0; for (int q = 1; q < 12; q++) if (cs.pool[q].d > cs.pool[wi].d) wi = q;
      if (d >= cs.pool[wi].d) continue;
      cs.np = 12;
      NpOut& o = cs.pool[wi];
      f3 on_v = v - nf * rV, on_f = proj + nf * rP;
      if (v_is_a) { o.pa = on_v; o.pb = on_f; o.n = nf; } else { o.pa = on_f; o.pb = on_v; o.n = -nf; }
      o.d = d;
      continue;
    }
    NpOut& o = cs.pool[cs.np++];
    f3 on_v = v - nf * rV, on_f = proj + nf * rP;
    if (v_is_a) { o.pa = on_v; o.pb = on_f; o.n = nf; } else { o.pa = on_f; o.pb = on_v; o.n = -nf; }
    o.d = d;
  }
}

// Manifold selection.  The GJK primary point is arbitrary within a flat contact patch (any point of
// two parallel faces is "closest"), so whenever feature candidates exist the manifold is built from
// them only: deepest candidate first, then greedily the candidate farthest from the chosen set.
AG_HDN inline int select_cands(const CandSel& cs, NpOut* out) {
  if (cs.np == 0) { out[0] = cs.prim; return 1; }
  bool used[12];
  int first = 0;
  for (int i = 0; i < 12; i++) used[i] = i >= cs.np;
  for (int i = 1; i < cs.np; i++) if (cs.pool[i].d < cs.pool[first].d) first = i;
  int nc = 0; out[nc++] = cs.pool[first]; used[first] = true;
  while (nc < 4) {
    int bi = -1; float bd = 1e-8f;
    for (int i = 0; i < cs.np; i++) {
      if (used[i]) continue;
      float md = 1e30f;
      for (int k = 0; k < nc; k++) { f3 d = cs.pool[i].pa - out[k].pa; md = fminf(md, dot(d, d)); }
      if (md > bd) { bd = md; bi = i; }
    }
    if (bi < 0) break;
    used[bi] = true; out[nc++] = cs.pool[bi];
  }
  return nc;
}

// contacts between colliders ca (A) and cb (B) of env e; results in WORLD coordinates.
AG_HDN inline int narrow_pair(const SimDev& S, int e, int ca, int cb, float max_dist, bool manifold, NpOut* out) {
  const int N = S.N;
  int ta = AG_LDG(S.col_type + ca), tb = AG_LDG(S.col_type + cb);
  float ra = AG_LDG(S.col_radius + ca), rb = AG_LDG(S.col_radius + cb);
  int ka = AG_LDG(S.col_link + ca), kb = AG_LDG(S.col_link + cb);
  int va0 = AG_LDG(S.col_v0 + ca), nA = AG_LDG(S.col_nv + ca), vb0 = AG_LDG(S.col_v0 + cb), nB = AG_LDG(S.col_nv + cb);
  int pa0 = AG_LDG(S.col_p0 + ca), npA = AG_LDG(S.col_np + ca), pb0 = AG_LDG(S.col_p0 + cb), npB = AG_LDG(S.col_np + cb);
  f3 posA = ld3(S.lpos, ka, N, e), posB = ld3(S.lpos, kb, N, e);
  q4 qA = ld4(S.lquat, ka, N, e), qB = ld4(S.lquat, kb, N, e);
  CandSel cs; cs.np = 0;
  int nout;
  if (ta == 3 || tb == 3) {
    if (ta == tb) return 0;
    bool flip = (ta == 3);                 // half-space is A; compute in the half-space owner's frame
    // work in the plane owner's local frame: treat owner as "B" of the local computation
    q4 qP = flip ? qA : qB, qV = flip ? qB : qA;
    f3 pP = flip ? posA : posB, pV = flip ? posB : posA;
    m3 R = mul(transpose(qmat(qP)), qmat(qV));
    f3 t = qrot_inv(qP, pV - pP);
    int vv0 = flip ? vb0 : va0, nV = flip ? nB : nA; float rv = flip ? rb : ra;
    f3 pn; float pd; ld_plane(S.planes, flip ? pa0 : pb0, pn, pd);
    int j = 0; float mn = 1e30f;
    for (int i = 0; i < nV; i++) { float x = dot(pn, mul(R, tv3(S.verts, vv0 + i)) + t); if (x < mn) { mn = x; j = i; } }
    float d = mn - pd - rv;
    if (d > max_dist) return 0;
    f3 vj = mul(R, tv3(S.verts, vv0 + j)) + t;
    // local result with V playing "A" (normal from plane towards V)
    cs.prim.pa = vj - pn * rv; cs.prim.pb = vj - pn * (mn - pd); cs.prim.n = pn; cs.prim.d = d;
    if (manifold && nV > 1) face_cands(S, vv0, nV, rv, true, flip ? pa0 : pb0, 1, 0.f, R, t, pn, d, max_dist * 0.5f, max_dist, cs);
    nout = select_cands(cs, out);
    m3 RP = qmat(qP);
    for (int i = 0; i < nout; i++) {
      f3 a = mul(RP, out[i].pa) + pP, b = mul(RP, out[i].pb) + pP, n = mul(RP, out[i].n);
      if (!flip) { out[i].pa = a; out[i].pb = b; out[i].n = n; }
      else { out[i].pa = b; out[i].pb = a; out[i].n = -n; }
    }
    return nout;
  }
  m3 RB = qmat(qB);
  m3 R = mul(transpose(RB), qmat(qA));
  f3 t = mulT(RB, posA - posB);
  f3 pa, pb, nrm; float dist = 0.f;
  bool ov = gjk_cores(S.verts, va0, nA, vb0, nB, R, t, pa, pb, nrm, dist);
  if (ov) pen_faces(S, va0, nA, pa0, npA, vb0, nB, pb0, npB, R, t, pa, pb, nrm, dist);
  float d = dist - ra - rb;
  if (d > max_dist) return 0;
  cs.prim.n = nrm; cs.prim.pa = pa - nrm * ra; cs.prim.pb = pb + nrm * rb; cs.prim.d = d;
  if (manifold) {
    if (npB > 0 && nA > 1) face_cands(S, va0, nA, ra, true, pb0, npB, rb, R, t, nrm, d, max_dist * 0.5f, max_dist, cs);
    if (npA > 0 && nB > 1) face_cands(S, vb0, nB, rb, false, pa0, npA, ra, R, t, -nrm, d, max_dist * 0.5f, max_dist, cs);
  }
  nout = select_cands(cs, out);
  for (int i = 0; i < nout; i++) {
    out[i].pa = mul(RB, out[i].pa) + posB; out[i].pb = mul(RB, out[i].pb) + posB; out[i].n = mul(RB, out[i].n);
  }
  return nout;
}

AG_HD int ag_atomic_inc(int* p) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, 1);
#else
  int v = *p; *p = v + 1; return v;
#endif
}
AG_HD int ag_atomic_add(int* p, int k) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, k);
#else
  int v = *p; *p = v + k; return v;
#endif
}

// K3a: thread = (link pair p, env lane), env fastest; p.i0 = padded env count.  Cheap AABB culls only:
// surviving collider pairs are appended to the env's candidate list.  Light kernel (few registers,
// full occupancy); the heavy GJK work runs in K3b with one thread per candidate.
AG_HDN inline void pairs_body(int tid, const SimDev& S, const KP& kp) {
  const int N = S.N;
  int Npad = kp.i0;
  int e = tid % Npad, pr = tid / Npad;
  if (e >= N) return;
  int la = AG_LDG(S.pair_link + 2 * pr), lb = AG_LDG(S.pair_link + 2 * pr + 1);
  int ba = AG_LDG(S.link_body + la), bb = AG_LDG(S.link_body + lb);
  if (S.body_mode[(size_t)ba * N + e] == 0 || S.body_mode[(size_t)bb * N + e] == 0) return;
  float fac = S.contact_thr;
  f3 lamin = ld3(S.lmin, la, N, e), lamax = ld3(S.lmax, la, N, e), lbmin = ld3(S.lmin, lb, N, e), lbmax = ld3(S.lmax, lb, N, e);
  float tla = AG_LDG(S.link_thresh + la), tlb = AG_LDG(S.link_thresh + lb);
  if (!aabb_ov(lamin, lamax, lbmin, lbmax, fac * fminf(tla, tlb))) return;
  int ca0 = AG_LDG(S.link_col0 + la), nca = AG_LDG(S.link_ncol + la), cb0 = AG_LDG(S.link_col0 + lb), ncb = AG_LDG(S.link_ncol + lb);
  for (int ca = ca0; ca < ca0 + nca; ca++) {
    f3 amin = ld3(S.cmin, ca, N, e), amax = ld3(S.cmax, ca, N, e);
    float tha = AG_LDG(S.col_thresh + ca);
    if (!aabb_ov(amin, amax, lbmin, lbmax, fac * fminf(tha, tlb))) continue;
    for (int cb = cb0; cb < cb0 + ncb; cb++) {
      float thr = fac * fminf(tha, AG_LDG(S.col_thresh + cb));   // size-relative breaking threshold
      if (!aabb_ov(amin, amax, ld3(S.cmin, cb, N, e), ld3(S.cmax, cb, N, e), thr)) continue;
      int slot = ag_atomic_inc(S.cand_count + e);
      if (slot < S.maxcand) S.cand[(size_t)slot * N + e] = (unsigned)ca * (unsigned)S.nc + (unsigned)cb;
    }
  }
}

// K3b: thread = (candidate slot, env): GJK / face fallback / manifold for one collider pair.
AG_HDN inline void narrow_body(int tid, const SimDev& S, const KP&) {
  const int N = S.N;
  int e = tid % N, cs = tid / N;
  int ncand = S.cand_count[e]; if (ncand > S.maxcand) ncand = S.maxcand;
  if (cs >= ncand) return;
  unsigned pk = S.cand[(size_t)cs * N + e];
  int ca = (int)(pk / (unsigned)S.nc), cb = (int)(pk % (unsigned)S.nc);
  float thr = S.contact_thr * fminf(AG_LDG(S.col_thresh + ca), AG_LDG(S.col_thresh + cb));
  NpOut out[4];
  int n = narrow_pair(S, e, ca, cb, thr, true, out);
  for (int i = 0; i < n; i++) {
    int slot = ag_atomic_inc(S.c_count + e);
    if (slot >= S.maxc) continue;
    S.c_key[(size_t)slot * N + e] = pk * 4u + (unsigned)i;
    cf_st(S.c_data, slot, CF_PAX, N, e, out[i].pa.x); cf_st(S.c_data, slot, CF_PAY, N, e, out[i].pa.y); cf_st(S.c_data, slot, CF_PAZ, N, e, out[i].pa.z);
    cf_st(S.c_data, slot, CF_PBX, N, e, out[i].pb.x); cf_st(S.c_data, slot, CF_PBY, N, e, out[i].pb.y); cf_st(S.c_data, slot, CF_PBZ, N, e, out[i].pb.z);
    cf_st(S.c_data, slot, CF_NX, N, e, out[i].n.x); cf_st(S.c_data, slot, CF_NY, N, e, out[i].n.y); cf_st(S.c_data, slot, CF_NZ, N, e, out[i].n.z);
    cf_st(S.c_data, slot, CF_DIST, N, e, out[i].d);
  }
}

// K4: deterministic order: rank each contact by its key.  thread = (slot, env).
AG_HDN inline void sort_body(int tid, const SimDev& S, const KP&) {
  const int N = S.N;
  int e = tid % N, slot = tid / N;
  int cnt = S.c_count[e];
  int n = cnt < S.maxc ? cnt : S.maxc;
  if (slot == 0) S.overflow[e] = (cnt > S.maxc) || (S.cand_count[e] > S.maxcand);
  if (slot >= n) return;
  unsigned key = S.c_key[(size_t)slot * N + e];
  int rank = 0;
  for (int j = 0; j < n; j++) rank += (S.c_key[(size_t)j * N + e] < key) ? 1 : 0;
  S.s_key[(size_t)rank * N + e] = key;
  for (int f = 0; f <= CF_DIST; f++) cf_st(S.s_data, rank, f, N, e, cf_ld(S.c_data, slot, f, N, e));
}

// ------------------------------------------------------------------ K5: unconstrained dynamics
// 6x6 articulated inertia in world-aligned axes about the link origin: [[A, B],[B^T, D]], A and D symmetric
struct AI { s3 A; m3 B; s3 D; };
struct SVf { f3 a, l; };   // spatial vector (angular, linear)

AG_HD SVf ai_mul(const AI& I, SVf v) {
  SVf r;
  r.a = mul(I.A, v.a) + mul(I.B, v.l);
  r.l = mulT(I.B, v.a) + mul(I.D, v.l);
  return r;
}
AG_HD m3 skew_m(f3 v) { m3 r; r.m[0] = 0; r.m[1] = -v.z; r.m[2] = v.y; r.m[3] = v.z; r.m[4] = 0; r.m[5] = -v.x; r.m[6] = -v.y; r.m[7] = v.x; r.m[8] = 0; return r; }
AG_HD m3 s3_to_m3(const s3& s) { m3 r; r.m[0] = s.xx; r.m[1] = s.xy; r.m[2] = s.xz; r.m[3] = s.xy; r.m[4] = s.yy; r.m[5] = s.yz; r.m[6] = s.xz; r.m[7] = s.yz; r.m[8] = s.zz; return r; }
// move the reference point of an inertia from P to O where P = O + r
AG_HD AI ai_shift(const AI& I, f3 r) {
  AI o;
  m3 rx = skew_m(r);
  m3 D = s3_to_m3(I.D);
  m3 rxD = mul(rx, D);                    // r x D
  o.D = I.D;
  for (int i = 0; i < 9; i++) o.B.m[i] = I.B.m[i] + rxD.m[i];
  // A_O = A + rx B^T - B rx - rx D rx
  m3 rxBT = mul(rx, transpose(I.B));
  m3 Brx = mul(I.B, rx);
  m3 rxDrx = mul(rxD, rx);
  m3 A = s3_to_m3(I.A);
  for (int i = 0; i < 9; i++) A.m[i] = A.m[i] + rxBT.m[i] - Brx.m[i] - rxDrx.m[i];
  o.A.xx = A.m[0]; o.A.yy = A.m[4]; o.A.zz = A.m[8];
  o.A.xy = 0.5f * (A.m[1] + A.m[3]); o.A.xz = 0.5f * (A.m[2] + A.m[6]); o.A.yz = 0.5f * (A.m[5] + A.m[7]);
  return o;
}
AG_HD void ai_add(AI& a, const AI& b) {
  a.A.xx += b.A.xx; a.A.yy += b.A.yy; a.A.zz += b.A.zz; a.A.xy += b.A.xy; a.A.xz += b.A.xz; a.A.yz += b.A.yz;
  for (int i = 0; i < 9; i++) a.B.m[i] += b.B.m[i];
  a.D.xx += b.D.xx; a.D.yy += b.D.yy; a.D.zz += b.D.zz; a.D.xy += b.D.xy; a.D.xz += b.D.xz; a.D.yz += b.D.yz;
}
// I - U U^T * s
AG_HD AI ai_sub_outer(const AI& I, SVf U, float s) {
  AI o = I;
  o.A.xx -= U.a.x * U.a.x * s; o.A.yy -= U.a.y * U.a.y * s; o.A.zz -= U.a.z * U.a.z * s;
  o.A.xy -= U.a.x * U.a.y * s; o.A.xz -= U.a.x * U.a.z * s; o.A.yz -= U.a.y * U.a.z * s;
  float ua[3] = {U.a.x, U.a.y, U.a.z}, ul[3] = {U.l.x, U.l.y, U.l.z};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o.B.m[3 * i + j] -= ua[i] * ul[j] * s;
  o.D.xx -= U.l.x * U.l.x * s; o.D.yy -= U.l.y * U.l.y * s; o.D.zz -= U.l.z * U.l.z * s;
  o.D.xy -= U.l.x * U.l.y * s; o.D.xz -= U.l.x * U.l.z * s; o.D.yz -= U.l.y * U.l.z * s;
  return o;
}
AG_HD SVf sv_add(SVf a, SVf b) { SVf r; r.a = a.a + b.a; r.l = a.l + b.l; return r; }
AG_HD SVf sv_scale(SVf a, float s) { SVf r; r.a = a.a * s; r.l = a.l * s; return r; }
AG_HD float sv_dot(SVf m, SVf f) { return dot(m.a, f.a) + dot(m.l, f.l); }

// One lane per env: free bodies (gravity, damping, gyroscopic) and articulated bodies (ABA + M^-1).
AG_HDN inline void dyn_body(int e, const SimDev& S, const KP&) {
  const int N = S.N;
  const float dt = S.dt, vmax = S.vmax, kl = S.lin_damp, ka = S.ang_damp;
  // ---- free rigid bodies
  for (int f = 0; f < S.nf; f++) {
    int b = AG_LDG(S.free_body + f);
    int l0 = AG_LDG(S.body_link0 + b);
    q4 q = ld4(S.lquat, l0, N, e);
    f3 com = ld3(S.lpos, l0, N, e) + qrot(q, tv3(S.link_com, l0));
    st3(S.fcom, f, N, e, com);
    size_t ib = (size_t)f * 6 * N + e;
    if (S.body_mode[(size_t)b * N + e] != 1) {
      for (int i = 0; i < 6; i++) S.fIinv[ib + (size_t)i * N] = 0.f;
      continue;
    }
    m3 R = qmat(qmul(q, tv4(S.link_iquat, l0)));
    f3 Id = tv3(S.link_inertia, l0);
    s3 Il; Il.xx = Id.x; Il.yy = Id.y; Il.zz = Id.z; Il.xy = Il.xz = Il.yz = 0.f;
    s3 Iw = rot_sym(R, Il);
    s3 Ii; Ii.xx = 1.0f / Id.x; Ii.yy = 1.0f / Id.y; Ii.zz = 1.0f / Id.z; Ii.xy = Ii.xz = Ii.yz = 0.f;
    s3 Iinv = rot_sym(R, Ii);
    f3 v = ld3(S.base_lin, b, N, e), w = ld3(S.base_ang, b, N, e);
    f3 g = tv3(S.body_gravity, b);
    f3 acc = g - v * (kl + kl * norm(v));
    f3 Iww = mul(Iw, w);
    f3 tau = -(Iww * (ka + ka * norm(w)));
    if (S.gyro) tau = tau - cross(w, Iww);
    v = v + acc * dt; w = w + mul(Iinv, tau) * dt;
    v = f3(clampf(v.x, -vmax, vmax), clampf(v.y, -vmax, vmax), clampf(v.z, -vmax, vmax));
    w = f3(clampf(w.x, -vmax, vmax), clampf(w.y, -vmax, vmax), clampf(w.z, -vmax, vmax));
    st3(S.base_lin, b, N, e, v); st3(S.base_ang, b, N, e, w);
    S.fIinv[ib] = Iinv.xx; S.fIinv[ib + N] = Iinv.yy; S.fIinv[ib + 2 * (size_t)N] = Iinv.zz;
    S.fIinv[ib + 3 * (size_t)N] = Iinv.xy; S.fIinv[ib + 4 * (size_t)N] = Iinv.xz; S.fIinv[ib + 5 * (size_t)N] = Iinv.yz;
  }
  // ---- articulated bodies
  for (int a = 0; a < S.nart; a++) {
    int b = AG_LDG(S.art_body + a), d0 = AG_LDG(S.art_dl0 + a), nd = AG_LDG(S.art_nd + a);
    bool active = S.body_mode[(size_t)b * N + e] == 1;
    AI IA[AG_MAXND]; SVf pA[AG_MAXND], U[AG_MAXND], c[AG_MAXND], vel[AG_MAXND];
    f3 ax[AG_MAXND], rr[AG_MAXND];
    float Dinv[AG_MAXND], u[AG_MAXND], qd[AG_MAXND];
    int par[AG_MAXND], typ[AG_MAXND];
    f3 g = tv3(S.body_gravity, b);
    int l0 = AG_LDG(S.body_link0 + b);
    f3 obase = ld3(S.lpos, l0, N, e);
    // pass 1: velocities, bias, rigid inertias (world-aligned axes, referred to each link's origin)
    for (int i = 0; i < nd; i++) {
      int d = d0 + i, k = AG_LDG(S.dl_link + d);
      par[i] = AG_LDG(S.dl_parent + d); if (par[i] >= 0) par[i] -= d0;
      typ[i] = AG_LDG(S.dl_type + d);
      q4 q = ld4(S.lquat, k, N, e);
      f3 o = ld3(S.lpos, k, N, e);
      m3 R = qmat(q);
      ax[i] = mul(R, tv3(S.link_axis, k));
      st3(S.jax, d, N, e, ax[i]); st3(S.jor, d, N, e, o);
      f3 op = par[i] >= 0 ? ld3(S.lpos, AG_LDG(S.dl_link + d0 + par[i]), N, e) : obase;
      rr[i] = o - op;
      qd[i] = active ? ld1(S.jqd, k, N, e) : 0.f;
      SVf vp; if (par[i] >= 0) vp = vel[par[i]];
      SVf v; v.a = vp.a; v.l = vp.l + cross(vp.a, rr[i]);
      SVf vj; if (typ[i] == 1) { vj.a = ax[i] * qd[i]; } else { vj.l = ax[i] * qd[i]; }
      v = sv_add(v, vj);
      vel[i] = v;
      c[i].a = cross(v.a, vj.a); c[i].l = cross(v.a, vj.l) + cross(v.l, vj.a);
      // rigid inertia about the link origin
      float m = AG_LDG(S.dl_mass + d);
      f3 mc = mul(R, tv3(S.dl_mc, d));
      s3 Jl; Jl.xx = AG_LDG(S.dl_J + 6 * d); Jl.yy = AG_LDG(S.dl_J + 6 * d + 1); Jl.zz = AG_LDG(S.dl_J + 6 * d + 2);
      Jl.xy = AG_LDG(S.dl_J + 6 * d + 3); Jl.xz = AG_LDG(S.dl_J + 6 * d + 4); Jl.yz = AG_LDG(S.dl_J + 6 * d + 5);
      IA[i].A = rot_sym(R, Jl);
      IA[i].B = skew_m(mc);
      IA[i].D.xx = IA[i].D.yy = IA[i].D.zz = m; IA[i].D.xy = IA[i].D.xz = IA[i].D.yz = 0.f;
      SVf Iv = ai_mul(IA[i], v);
      pA[i].a = cross(v.a, Iv.a) + cross(v.l, Iv.l);
      pA[i].l = cross(v.a, Iv.l);
      // velocity damping, applied per rigid part at its COM (Bullet applies it per original link)
      int p0 = AG_LDG(S.dl_part0 + d), npt = AG_LDG(S.dl_nparts + d);
      for (int pi = p0; pi < p0 + npt; pi++) {
        float pm = AG_LDG(S.pt_mass + pi);
        f3 pc = mul(R, tv3(S.pt_com, pi));
        f3 vc = v.l + cross(v.a, pc);
        f3 fd = vc * (-pm * (kl + kl * norm(vc)));
        s3 Ip; Ip.xx = AG_LDG(S.pt_I + 6 * pi); Ip.yy = AG_LDG(S.pt_I + 6 * pi + 1); Ip.zz = AG_LDG(S.pt_I + 6 * pi + 2);
        Ip.xy = AG_LDG(S.pt_I + 6 * pi + 3); Ip.xz = AG_LDG(S.pt_I + 6 * pi + 4); Ip.yz = AG_LDG(S.pt_I + 6 * pi + 5);
        f3 nd_ = mul(rot_sym(R, Ip), v.a) * (-(ka + ka * norm(v.a)));
        pA[i].a = pA[i].a - (nd_ + cross(pc, fd));
        pA[i].l = pA[i].l - fd;
      }
    }
    // pass 2: articulated inertias, leaf to root
    for (int i = nd - 1; i >= 0; i--) {
      SVf Sx; if (typ[i] == 1) Sx.a = ax[i]; else Sx.l = ax[i];
      U[i] = ai_mul(IA[i], Sx);
      float D = sv_dot(Sx, U[i]);
      Dinv[i] = 1.0f / D;
      float tau = -AG_LDG(S.dl_damping + d0 + i) * qd[i];
      u[i] = tau - sv_dot(Sx, pA[i]);
      if (par[i] >= 0) {
        AI Ia = ai_sub_outer(IA[i], U[i], Dinv[i]);
        SVf pa = sv_add(sv_add(pA[i], ai_mul(Ia, c[i])), sv_scale(U[i], u[i] * Dinv[i]));
        AI Is = ai_shift(Ia, rr[i]);
        ai_add(IA[par[i]], Is);
        pA[par[i]].a = pA[par[i]].a + pa.a + cross(rr[i], pa.l);
        pA[par[i]].l = pA[par[i]].l + pa.l;
      }
    }
    // pass 3: accelerations, root to leaf
    SVf acc[AG_MAXND];
    for (int i = 0; i < nd; i++) {
      SVf ap; if (par[i] >= 0) ap = acc[par[i]]; else { ap.l = -g; }
      SVf a1; a1.a = ap.a + c[i].a; a1.l = ap.l + cross(ap.a, rr[i]) + c[i].l;
      float qdd = (u[i] - sv_dot(U[i], a1)) * Dinv[i];
      if (typ[i] == 1) a1.a = a1.a + ax[i] * qdd; else a1.l = a1.l + ax[i] * qdd;
      acc[i] = a1;
      float nq = clampf(qd[i] + dt * qdd, -vmax, vmax);
      if (active) st1(S.jqd, AG_LDG(S.dl_link + d0 + i), N, e, nq);
    }
    // M^-1 by unit joint impulses through the cached articulated inertias
    for (int j = 0; j < nd; j++) {
      SVf p[AG_MAXND]; float uu[AG_MAXND];
      for (int i = 0; i < nd; i++) { p[i] = SVf(); }
      for (int i = nd - 1; i >= 0; i--) {
        SVf Sx; if (typ[i] == 1) Sx.a = ax[i]; else Sx.l = ax[i];
        uu[i] = ((i == j) ? 1.f : 0.f) - sv_dot(Sx, p[i]);
        if (par[i] >= 0) {
          SVf pa = sv_add(p[i], sv_scale(U[i], uu[i] * Dinv[i]));
          p[par[i]].a = p[par[i]].a + pa.a + cross(rr[i], pa.l);
          p[par[i]].l = p[par[i]].l + pa.l;
        }
      }
      SVf aa[AG_MAXND];
      for (int i = 0; i < nd; i++) {
        SVf ap; if (par[i] >= 0) ap = aa[par[i]];
        SVf a1; a1.a = ap.a; a1.l = ap.l + cross(ap.a, rr[i]);
        float qdd = (uu[i] - sv_dot(U[i], a1)) * Dinv[i];
        if (typ[i] == 1) a1.a = a1.a + ax[i] * qdd; else a1.l = a1.l + ax[i] * qdd;
        aa[i] = a1;
        S.Minv[((size_t)(d0 + i) * S.ND + (d0 + j)) * N + e] = active ? qdd : 0.f;
      }
    }
  }
}

// ------------------------------------------------------------------ K6: constraint rows
// side reference encoding: (idx << 2) | kind, kind: 0 static, 1 free body (idx = f), 2 articulated (idx = dyn link)
AG_HD int link_ref(const SimDev& S, int e, int link) {
  int b = AG_LDG(S.link_body + link);
  int kind = AG_LDG(S.body_kind + b);
  if (S.body_mode[(size_t)b * S.N + e] != 1) return 0;
  if (kind == BK_FREE) return (AG_LDG(S.body_idx + b) << 2) | 1;
  if (kind == BK_ART) { int d = AG_LDG(S.link_dl + link); return d < 0 ? 0 : ((d << 2) | 2); }
  return 0;
}

AG_HD s3 ld_Iinv(const SimDev& S, int f, int e) {
  size_t ib = (size_t)f * 6 * S.N + e; size_t N = S.N;
  s3 r; r.xx = S.fIinv[ib]; r.yy = S.fIinv[ib + N]; r.zz = S.fIinv[ib + 2 * N]; r.xy = S.fIinv[ib + 3 * N]; r.xz = S.fIinv[ib + 4 * N]; r.yz = S.fIinv[ib + 5 * N];
  return r;
}

// Fill the articulated side slot `as` with J (unit force `lin` at world point p + torque `ang` on dyn link d)
// and M^-1 J^T; returns J M^-1 J^T and accumulates J.qd into rel.
AG_HDN inline float art_side(const SimDev& S, int e, int as, int d, f3 p, f3 lin, f3 ang, float& rel) {
  const int N = S.N;
  float J[AG_MAXND];
  int a = AG_LDG(S.dl_art + d), d0 = AG_LDG(S.art_dl0 + a), nd = AG_LDG(S.art_nd + a);
  for (int i = 0; i < nd; i++) J[i] = 0.f;
  int j = d;
  while (j >= 0) {
    f3 axw = ld3(S.jax, j, N, e), o = ld3(S.jor, j, N, e);
    J[j - d0] = (AG_LDG(S.dl_type + j) == 1) ? (dot(lin, cross(axw, p - o)) + dot(ang, axw)) : dot(lin, axw);
    j = AG_LDG(S.dl_parent + j);
  }
  float diag = 0.f;
  for (int i = 0; i < nd; i++) {
    float m = 0.f;
    for (int k = 0; k < nd; k++) m += S.Minv[((size_t)(d0 + i) * S.ND + (d0 + k)) * N + e] * J[k];
    S.as_J[((size_t)as * AG_MAXND + i) * N + e] = J[i];
    S.as_MiJ[((size_t)as * AG_MAXND + i) * N + e] = m;
    diag += J[i] * m;
    rel += J[i] * ld1(S.jqd, AG_LDG(S.dl_link + d0 + i), N, e);
  }
  return diag;
}

// diag and relative velocity contribution of one side for direction `dir` (force on this side = sign*dir at p)
AG_HDN inline float side_terms(const SimDev& S, int e, int ref, int as, f3 p, f3 lin, f3 ang, float& rel) {
  int kind = ref & 3, idx = ref >> 2;
  if (kind == 1) {
    int b = AG_LDG(S.free_body + idx);
    float invm = 1.0f / AG_LDG(S.link_mass + AG_LDG(S.body_link0 + b));
    f3 r = p - ld3(S.fcom, idx, S.N, e);
    f3 t = cross(r, lin) + ang;
    s3 Ii = ld_Iinv(S, idx, e);
    f3 v = ld3(S.base_lin, b, S.N, e), w = ld3(S.base_ang, b, S.N, e);
    rel += dot(lin, v) + dot(t, w);
    return invm * dot(lin, lin) + dot(t, mul(Ii, t));
  } else if (kind == 2) {
    return art_side(S, e, as, idx, p, lin, ang, rel);
  }
  return 0.f;
}

// K6a: one lane per env: joint-limit rows, motor rows, fixed-constraint rows
AG_HDN inline void rows_body(int e, const SimDev& S, const KP&) {
  const int N = S.N;
  const float dt = S.dt;
  S.as_count[e] = 6 * S.ncon;   // art-side slots [0, 6 ncon) are reserved for side A/B of the fixed constraints
  for (int d = 0; d < S.ND; d++) {
    int k = AG_LDG(S.dl_link + d);
    float Mdd = S.Minv[((size_t)d * S.ND + d) * N + e];
    float q = ld1(S.jq, k, N, e), qd = ld1(S.jqd, k, N, e);
    float dinv = Mdd > 0.f ? 1.0f / Mdd : 0.f;
    // limits: a row only while violated
    float rl = 0.f, dl = 0.f, ru = 0.f, du = 0.f;
    if (AG_LDG(S.link_haslimit + k) && dinv > 0.f) {
      float penl = q - AG_LDG(S.link_lower + k), penu = AG_LDG(S.link_upper + k) - q;
      if (penl <= 0.f) { dl = dinv; rl = (-penl * S.erp / dt - qd) * dinv; }
      if (penu <= 0.f) { du = dinv; ru = (-penu * S.erp / dt + qd) * dinv; }
    }
    st1(S.dr_rhs, d, N, e, rl); st1(S.dr_dinv, d, N, e, dl); st1(S.dr_lam, d, N, e, 0.f);
    st1(S.dr_rhs, S.ND + d, N, e, ru); st1(S.dr_dinv, S.ND + d, N, e, du); st1(S.dr_lam, S.ND + d, N, e, 0.f);
    // motor
    float rm = 0.f, dm = 0.f;
    int mode = S.motor_mode[k];
    float maxi = S.motor_maxf[k] * dt;
    if (mode != 0 && maxi > 0.f && dinv > 0.f) {
      float vt = (mode == 1) ? (S.motor_kp[k] * (ld1(S.motor_target, k, N, e) - q) / dt + qd - S.motor_kd[k] * qd)
                             : ld1(S.motor_target, k, N, e);
      dm = dinv; rm = (vt - qd) * dinv;
    }
    st1(S.dr_rhs, 2 * S.ND + d, N, e, rm); st1(S.dr_dinv, 2 * S.ND + d, N, e, dm); st1(S.dr_lam, 2 * S.ND + d, N, e, 0.f);
  }
  for (int c = 0; c < S.ncon; c++) {
    int ka = AG_LDG(S.con_link + 2 * c), kb = AG_LDG(S.con_link + 2 * c + 1);
    int refA = link_ref(S, e, ka), refB = link_ref(S, e, kb);
    int ba = AG_LDG(S.link_body + ka), bb = AG_LDG(S.link_body + kb);
    bool on = S.body_mode[(size_t)ba * N + e] != 0 && S.body_mode[(size_t)bb * N + e] != 0;
    q4 qa = ld4(S.lquat, ka, N, e), qb = ld4(S.lquat, kb, N, e);
    f3 pa = ld3(S.lpos, ka, N, e) + qrot(qa, tv3(S.con_pivot, 2 * c));
    f3 pb = ld3(S.lpos, kb, N, e) + qrot(qb, tv3(S.con_pivot, 2 * c + 1));
    q4 fa = qmul(qa, tv4(S.con_quat, 2 * c)), fb = qmul(qb, tv4(S.con_quat, 2 * c + 1));
    q4 qe = qmul(fa, qconj(fb));
    if (qe.w < 0.f) qe = q4(-qe.x, -qe.y, -qe.z, -qe.w);
    f3 perr = pa - pb, aerr(2.f * qe.x, 2.f * qe.y, 2.f * qe.z);
    float maxi = AG_LDG(S.con_maxforce + c) * dt;
    for (int i = 0; i < 6; i++) {
      int r = 6 * c + i;
      f3 axv(i % 3 == 0 ? 1.f : 0.f, i % 3 == 1 ? 1.f : 0.f, i % 3 == 2 ? 1.f : 0.f);
      f3 lin = i < 3 ? axv : f3(), ang = i < 3 ? f3() : axv;
      float rel = 0.f, diag = 0.f;
      // art-side slots: side A uses slot r (two-sided articulated constraints share: B uses 6 ncon + ... not supported)
      diag += side_terms(S, e, refA, r, pa, lin, ang, rel);
      float relb = 0.f;
      int asB = -1;
      if ((refB & 3) == 2) { asB = ag_atomic_inc(S.as_count + e); if (asB >= S.nas) { asB = -1; } }
      if ((refB & 3) != 2 || asB >= 0) diag += side_terms(S, e, refB, asB, pb, -lin, -ang, relb);
      rel += relb;
      float err = i < 3 ? comp(perr, i) : comp(aerr, i - 3);
      float dinv = (on && diag > 1e-20f) ? 1.0f / diag : 0.f;
      size_t gb = (size_t)r * 16 * N + e;
      f3 angA = f3(), angB = f3();
      if ((refA & 3) == 1) { angA = cross(pa - ld3(S.fcom, refA >> 2, N, e), lin) + ang; }
      if ((refB & 3) == 1) { angB = cross(pb - ld3(S.fcom, refB >> 2, N, e), lin) + ang; }
      S.gr_data[gb + (size_t)GR_LX * N] = lin.x; S.gr_data[gb + (size_t)GR_LY * N] = lin.y; S.gr_data[gb + (size_t)GR_LZ * N] = lin.z;
      S.gr_data[gb + (size_t)GR_AAX * N] = angA.x; S.gr_data[gb + (size_t)GR_AAY * N] = angA.y; S.gr_data[gb + (size_t)GR_AAZ * N] = angA.z;
      S.gr_data[gb + (size_t)GR_ABX * N] = angB.x; S.gr_data[gb + (size_t)GR_ABY * N] = angB.y; S.gr_data[gb + (size_t)GR_ABZ * N] = angB.z;
      S.gr_data[gb + (size_t)GR_RHS * N] = (-err * S.erp / dt - rel) * dinv;
      S.gr_data[gb + (size_t)GR_DINV * N] = dinv;
      S.gr_data[gb + (size_t)GR_LO * N] = -maxi; S.gr_data[gb + (size_t)GR_HI * N] = maxi;
      S.gr_data[gb + (size_t)GR_LAM * N] = 0.f;
      size_t rb = (size_t)r * 4 * N + e;
      S.gr_ref[rb] = refA; S.gr_ref[rb + N] = refB; S.gr_ref[rb + 2 * (size_t)N] = ((refA & 3) == 2) ? r : -1; S.gr_ref[rb + 3 * (size_t)N] = asB;
    }
  }
}

// K6b: contact rows, thread = (sorted slot, env)
AG_HDN inline void crows_body(int tid, const SimDev& S, const KP&) {
  const int N = S.N;
  int e = tid % N, slot = tid / N;
  int cnt = S.c_count[e]; if (cnt > S.maxc) cnt = S.maxc;
  if (slot >= cnt) return;
  unsigned key = S.s_key[(size_t)slot * N + e];
  unsigned pairk = key >> 2;
  int ca = (int)(pairk / (unsigned)S.nc), cb = (int)(pairk % (unsigned)S.nc);
  int ka = AG_LDG(S.col_link + ca), kb = AG_LDG(S.col_link + cb);
  int refA = link_ref(S, e, ka), refB = link_ref(S, e, kb);
  f3 pa(cf_ld(S.s_data, slot, CF_PAX, N, e), cf_ld(S.s_data, slot, CF_PAY, N, e), cf_ld(S.s_data, slot, CF_PAZ, N, e));
  f3 pb(cf_ld(S.s_data, slot, CF_PBX, N, e), cf_ld(S.s_data, slot, CF_PBY, N, e), cf_ld(S.s_data, slot, CF_PBZ, N, e));
  f3 n(cf_ld(S.s_data, slot, CF_NX, N, e), cf_ld(S.s_data, slot, CF_NY, N, e), cf_ld(S.s_data, slot, CF_NZ, N, e));
  float dist = cf_ld(S.s_data, slot, CF_DIST, N, e);
  int asA = -1, asB = -1;
  if ((refA & 3) == 2) { asA = ag_atomic_add(S.as_count + e, 3); if (asA + 3 > S.nas) { asA = -1; refA = 0; } }
  if ((refB & 3) == 2) { asB = ag_atomic_add(S.as_count + e, 3); if (asB + 3 > S.nas) { asB = -1; refB = 0; } }
  f3 t1, t2; plane_space(n, t1, t2);
  float mu = ld1(S.friction, ka, N, e) * ld1(S.friction, kb, N, e);
  float dt = S.dt;
  for (int d = 0; d < 3; d++) {
    f3 dir = d == 0 ? n : (d == 1 ? t1 : t2);
    float rel = 0.f, diag = 0.f;
    diag += side_terms(S, e, refA, asA + d, pa, dir, f3(), rel);
    diag += side_terms(S, e, refB, asB + d, pb, -dir, f3(), rel);
    float dinv = diag > 1e-20f ? 1.0f / diag : 0.f;
    float rhs;
    if (d == 0) {
      float pen = dist + S.slop;
      float poserr, velerr = -rel;
      if (pen > 0.f) { poserr = 0.f; velerr -= pen / dt; } else poserr = -pen * S.contact_erp / dt;
      rhs = (poserr + velerr) * dinv;
    } else rhs = -rel * dinv;
    cf_st(S.s_data, slot, CF_RHS_N + 2 * d, N, e, rhs);
    cf_st(S.s_data, slot, CF_DINV_N + 2 * d, N, e, dinv);
    cf_st(S.s_data, slot, CF_LAM_N + d, N, e, 0.f);
  }
  cf_st(S.s_data, slot, CF_MU, N, e, mu);
  size_t rb = (size_t)slot * 4 * N + e;
  S.s_ref[rb] = refA; S.s_ref[rb + N] = refB; S.s_ref[rb + 2 * (size_t)N] = asA; S.s_ref[rb + 3 * (size_t)N] = asB;
}

// ------------------------------------------------------------------ K7: PGS
// One env per lane, a few lanes per CTA (AG_PGS_LANES, default 4).  The Gauss-Seidel chain of one env
// is strictly sequential, so the kernel is latency-bound per row; measured on B200 (ncu, round 1):
// with row constants read from global memory every iteration a row cost ~2 200 cycles (serialised L2
// round trips, 17 % L1 hit rate).  Therefore EVERYTHING the sweep touches is staged once into shared
// memory (`sm`, lane-strided => conflict-free): velocity deltas, per-body inverse inertias / COMs,
// the articulated M^-1, all row constants (contacts, dof rows, fixed-constraint rows incl. their
// articulated Jacobian sides) and all impulses.  Only the rare articulated sides of contact rows stay
// in global memory.  Layout (floats per lane):
//   [dv: ND+6nf][fcom: 3nf][fIinv: 6nf][invm: nf][Minv: ND*ND][lam: 3*maxc][dr: 5*(3ND)][gr: ngr*(16+2ND)][crec: 20*maxc][art sides: nas*2ND]
#define PGS_CREC 20
struct PgsLayout { int o_dv, o_fc, o_fi, o_fm, o_mi, o_lam, o_dr, o_gr, o_cr, o_as, total; };
AG_HD PgsLayout pgs_layout(const SimDev& S) {
  PgsLayout L;
  L.o_dv = 0; L.o_fc = L.o_dv + S.ND + 6 * S.nf; L.o_fi = L.o_fc + 3 * S.nf; L.o_fm = L.o_fi + 6 * S.nf; L.o_mi = L.o_fm + S.nf;
  L.o_lam = L.o_mi + S.ND * S.ND; L.o_dr = L.o_lam + 3 * S.maxc; L.o_gr = L.o_dr + 15 * S.ND;
  L.o_cr = L.o_gr + S.ngr * (16 + 2 * S.ND); L.o_as = L.o_cr + PGS_CREC * S.maxc; L.total = L.o_as + S.nas * 2 * S.ND;
  return L;
}
#define SMF(i) sm[(i) * LANES]
AG_HD float i2f_bits(int v) { float f; memcpy(&f, &v, 4); return f; }
AG_HD int f2i_bits(float f) { int v; memcpy(&v, &f, 4); return v; }

// J.dv of one side.  `artJ` >= 0: offset in sm of this side's articulated Jacobian (else global slot `as`)
template <int LANES>
AG_HD float pgs_side_jv(const SimDev& S, int e, const float* sm, const PgsLayout& L, int ref, int as, int artJ, f3 lin, f3 ang_free) {
  int kind = ref & 3, idx = ref >> 2;
  if (kind == 1) {
    int o = L.o_dv + S.ND + 6 * idx;
    return lin.x * SMF(o) + lin.y * SMF(o + 1) + lin.z * SMF(o + 2) + ang_free.x * SMF(o + 3) + ang_free.y * SMF(o + 4) + ang_free.z * SMF(o + 5);
  } else if (kind == 2) {
    int a = AG_LDG(S.dl_art + idx), d0 = AG_LDG(S.art_dl0 + a), nd = AG_LDG(S.art_nd + a);
    float t = 0.f;
    if (artJ < 0) artJ = L.o_as + as * 2 * S.ND;          // staged copy of as_J[as]
    for (int i = 0; i < nd; i++) t += SMF(artJ + i) * SMF(L.o_dv + d0 + i);
    return t;
  }
  return 0.f;
}
template <int LANES>
AG_HD void pgs_side_apply(const SimDev& S, int e, float* sm, const PgsLayout& L, int ref, int as, int artM, f3 lin, f3 ang_free, float dl) {
  int kind = ref & 3, idx = ref >> 2;
  if (kind == 1) {
    int o = L.o_dv + S.ND + 6 * idx;
    float invm = SMF(L.o_fm + idx) * dl;
    int fi = L.o_fi + 6 * idx;
    s3 Ii; Ii.xx = SMF(fi); Ii.yy = SMF(fi + 1); Ii.zz = SMF(fi + 2); Ii.xy = SMF(fi + 3); Ii.xz = SMF(fi + 4); Ii.yz = SMF(fi + 5);
    f3 ia = mul(Ii, ang_free);
    SMF(o) += lin.x * invm; SMF(o + 1) += lin.y * invm; SMF(o + 2) += lin.z * invm;
    SMF(o + 3) += ia.x * dl; SMF(o + 4) += ia.y * dl; SMF(o + 5) += ia.z * dl;
  } else if (kind == 2) {
    int a = AG_LDG(S.dl_art + idx), d0 = AG_LDG(S.art_dl0 + a), nd = AG_LDG(S.art_nd + a);
    if (artM < 0) artM = L.o_as + as * 2 * S.ND + S.ND;   // staged copy of as_MiJ[as]
    for (int i = 0; i < nd; i++) SMF(L.o_dv + d0 + i) += SMF(artM + i) * dl;
  }
}

// K6c: longest-processing-time-first order for K7.  The PGS chain of an env is sequential and its
// length varies 10x between envs (iterations used x rows), so CTAs are issued heaviest-first and envs
// of similar weight share a warp.  Work is predicted from this substep's contact count and the
// previous substep's iteration count.  64-bucket counting sort; p.p1 = histogram[64] (zeroed).
AG_HD int pgs_work_bucket(const SimDev& S, int e) {
  int cnt = S.c_count[e]; if (cnt > S.maxc) cnt = S.maxc;
  int it = S.iters_used[e]; if (it < 1) it = 1;
  int w = it * (3 * cnt + 3 * S.ND + S.ngr);          // predicted row updates
  int b = 63 - w / 320;                               // heaviest work -> bucket 0
  return b < 0 ? 0 : b;
}
AG_HDN inline void order_hist_body(int e, const SimDev& S, const KP& p) {
  ag_atomic_inc((int*)p.p1 + pgs_work_bucket(S, e));
}
AG_HDN inline void order_scatter_body(int e, const SimDev& S, const KP& p) {
  // p.p1 = exclusive prefix of the histogram (consumed by atomic increments)
  int pos = ag_atomic_inc((int*)p.p1 + pgs_work_bucket(S, e));
  S.pgs_order[pos] = e;
}
AG_HDN inline void order_prefix_body(int tid, const SimDev&, const KP& p) {
  if (tid != 0) return;
  int* h = (int*)p.p1; int acc = 0;
  for (int b = 0; b < 64; b++) { int c = h[b]; h[b] = acc; acc += c; }
}

template <int LANES>
AG_HDN inline void pgs_body(int slot, const SimDev& S, const KP&, float* sm) {
  const int e = S.pgs_order[slot];
  const int N = S.N;
  const int ND = S.ND;
  const PgsLayout L = pgs_layout(S);
  const int nvel = ND + 6 * S.nf;
#if defined(__CUDA_ARCH__)
  long long t_begin = clock64();
#endif
  // ---- stage the per-env solver state and all row constants into shared memory
  for (int i = 0; i < nvel; i++) SMF(L.o_dv + i) = 0.f;
  for (int i = 0; i < 3 * S.nf; i++) SMF(L.o_fc + i) = S.fcom[(size_t)i * N + e];
  for (int i = 0; i < 6 * S.nf; i++) SMF(L.o_fi + i) = S.fIinv[(size_t)i * N + e];
  for (int i = 0; i < S.nf; i++) SMF(L.o_fm + i) = AG_LDG(S.free_invm + i);
  for (int i = 0; i < ND * ND; i++) SMF(L.o_mi + i) = S.Minv[(size_t)i * N + e];
  int cnt = S.c_count[e]; if (cnt > S.maxc) cnt = S.maxc;
  for (int i = 0; i < 3 * cnt; i++) SMF(L.o_lam + i) = 0.f;
  for (int r = 0; r < 3 * ND; r++) {      // dof rows: lambda, rhs, dinv, lo, hi
    int d = r % ND;
    float hi = (r / ND == 2) ? S.motor_maxf[AG_LDG(S.dl_link + d)] * S.dt : 1e30f;
    SMF(L.o_dr + 5 * r) = 0.f; SMF(L.o_dr + 5 * r + 1) = S.dr_rhs[(size_t)r * N + e]; SMF(L.o_dr + 5 * r + 2) = S.dr_dinv[(size_t)r * N + e];
    SMF(L.o_dr + 5 * r + 3) = (r / ND == 2) ? -hi : 0.f; SMF(L.o_dr + 5 * r + 4) = hi;
  }
  const int GRW = 16 + 2 * ND;
  for (int r = 0; r < S.ngr; r++) {       // fixed-constraint rows: 16 fields (GR_LAM reused as lambda, PAD0/1 = refs) + art sides
    int o = L.o_gr + r * GRW;
    const float* g = S.gr_data + (size_t)r * 16 * N + e;
    for (int f = 0; f < 14; f++) SMF(o + f) = g[(size_t)f * N];
    SMF(o + GR_LAM) = 0.f;
    const int* rf = S.gr_ref + (size_t)r * 4 * N + e;
    int refA = rf[0], refB = rf[N], asA = rf[2 * (size_t)N], asB = rf[3 * (size_t)N];
    SMF(o + GR_PAD0) = i2f_bits(refA); SMF(o + GR_PAD1) = i2f_bits(refB);
    // articulated side (at most one side of a fixed constraint is staged; a second one falls back to global)
    int as = (refA & 3) == 2 ? asA : ((refB & 3) == 2 ? asB : -1);
    for (int i = 0; i < ND; i++) {
      SMF(o + 16 + i) = as >= 0 ? S.as_J[((size_t)as * AG_MAXND + i) * N + e] : 0.f;
      SMF(o + 16 + ND + i) = as >= 0 ? S.as_MiJ[((size_t)as * AG_MAXND + i) * N + e] : 0.f;
    }
  }
  for (int s = 0; s < cnt; s++) {         // contact records
    int o = L.o_cr + s * PGS_CREC;
    const float* c = S.s_data + (size_t)s * AG_CF * N + e;
    const int* rf = S.s_ref + (size_t)s * 4 * N + e;
    int refA = rf[0], refB = rf[N], asA = rf[2 * (size_t)N], asB = rf[3 * (size_t)N];
    f3 pa(c[(size_t)CF_PAX * N], c[(size_t)CF_PAY * N], c[(size_t)CF_PAZ * N]);
    f3 pb(c[(size_t)CF_PBX * N], c[(size_t)CF_PBY * N], c[(size_t)CF_PBZ * N]);
    f3 rA, rB;
    if ((refA & 3) == 1) { int q = L.o_fc + 3 * (refA >> 2); rA = f3(pa.x - SMF(q), pa.y - SMF(q + 1), pa.z - SMF(q + 2)); }
    if ((refB & 3) == 1) { int q = L.o_fc + 3 * (refB >> 2); rB = f3(pb.x - SMF(q), pb.y - SMF(q + 1), pb.z - SMF(q + 2)); }
    SMF(o) = c[(size_t)CF_NX * N]; SMF(o + 1) = c[(size_t)CF_NY * N]; SMF(o + 2) = c[(size_t)CF_NZ * N];
    SMF(o + 3) = rA.x; SMF(o + 4) = rA.y; SMF(o + 5) = rA.z; SMF(o + 6) = rB.x; SMF(o + 7) = rB.y; SMF(o + 8) = rB.z;
    SMF(o + 9) = c[(size_t)CF_RHS_N * N]; SMF(o + 10) = c[(size_t)CF_DINV_N * N];
    SMF(o + 11) = c[(size_t)CF_RHS_T1 * N]; SMF(o + 12) = c[(size_t)CF_DINV_T1 * N];
    SMF(o + 13) = c[(size_t)CF_RHS_T2 * N]; SMF(o + 14) = c[(size_t)CF_DINV_T2 * N];
    SMF(o + 15) = c[(size_t)CF_MU * N];
    SMF(o + 16) = i2f_bits(refA); SMF(o + 17) = i2f_bits(refB); SMF(o + 18) = i2f_bits(asA); SMF(o + 19) = i2f_bits(asB);
  }
  {                                        // articulated row sides allocated by k_rows / k_crows
    int nas_used = S.as_count[e]; if (nas_used > S.nas) nas_used = S.nas;
    for (int a = 0; a < nas_used; a++)
      for (int i = 0; i < ND; i++) {
        SMF(L.o_as + a * 2 * ND + i) = S.as_J[((size_t)a * AG_MAXND + i) * N + e];
        SMF(L.o_as + a * 2 * ND + ND + i) = S.as_MiJ[((size_t)a * AG_MAXND + i) * N + e];
      }
  }
  int used = 0;
  bool done = false;
#if defined(__CUDA_ARCH__)
  const unsigned wmask = __activemask();
#endif
  for (int it = 0; it < S.iters; it++) {
    if (!done) {
      float resid = 0.f;
      used = it + 1;
      // joint limits (lower, upper) then motors: J = +-e_d
      for (int r = 0; r < 3 * ND; r++) {
        float dinv = SMF(L.o_dr + 5 * r + 2);
        if (dinv == 0.f) continue;
        int d = r % ND; int kindr = r / ND;
        float sgn = kindr == 1 ? -1.f : 1.f;
        float lam = SMF(L.o_dr + 5 * r);
        float dl = SMF(L.o_dr + 5 * r + 1) - sgn * SMF(L.o_dv + d) * dinv;
        float lo = SMF(L.o_dr + 5 * r + 3), hi = SMF(L.o_dr + 5 * r + 4);
        float sum = lam + dl;
        if (sum < lo) { dl = lo - lam; sum = lo; } else if (sum > hi) { dl = hi - lam; sum = hi; }
        SMF(L.o_dr + 5 * r) = sum;
        int a = AG_LDG(S.dl_art + d), d0 = AG_LDG(S.art_dl0 + a), nd = AG_LDG(S.art_nd + a);
        float sdl = sgn * dl;
        for (int i = 0; i < nd; i++) SMF(L.o_dv + d0 + i) += SMF(L.o_mi + (d0 + i) * ND + d) * sdl;
        resid = fmaxf(resid, dl * dl);
      }
      // fixed constraints
      for (int r = 0; r < S.ngr; r++) {
        int o = L.o_gr + r * GRW;
        float dinv = SMF(o + GR_DINV);
        if (dinv == 0.f) continue;
        int refA = f2i_bits(SMF(o + GR_PAD0)), refB = f2i_bits(SMF(o + GR_PAD1));
        f3 lin(SMF(o + GR_LX), SMF(o + GR_LY), SMF(o + GR_LZ));
        f3 aA(SMF(o + GR_AAX), SMF(o + GR_AAY), SMF(o + GR_AAZ)), aB(SMF(o + GR_ABX), SMF(o + GR_ABY), SMF(o + GR_ABZ));
        bool aArt = (refA & 3) == 2;      // which side owns the staged articulated rows
        const int* rf = S.gr_ref + (size_t)r * 4 * N + e;
        int asB = (!aArt || (refB & 3) != 2) ? -1 : AG_LDG(rf + 3 * (size_t)N);
        float jv = pgs_side_jv<LANES>(S, e, sm, L, refA, -1, aArt ? o + 16 : -1, lin, aA) +
                   pgs_side_jv<LANES>(S, e, sm, L, refB, asB, (!aArt && (refB & 3) == 2) ? o + 16 : -1, -lin, -aB);
        float lam = SMF(o + GR_LAM);
        float dl = SMF(o + GR_RHS) - jv * dinv;
        float lo = SMF(o + GR_LO), hi = SMF(o + GR_HI);
        float sum = lam + dl;
        if (sum < lo) { dl = lo - lam; sum = lo; } else if (sum > hi) { dl = hi - lam; sum = hi; }
        SMF(o + GR_LAM) = sum;
        pgs_side_apply<LANES>(S, e, sm, L, refA, -1, aArt ? o + 16 + ND : -1, lin, aA, dl);
        pgs_side_apply<LANES>(S, e, sm, L, refB, asB, (!aArt && (refB & 3) == 2) ? o + 16 + ND : -1, -lin, -aB, dl);
        resid = fmaxf(resid, dl * dl);
      }
      // contact normals.  Free-body sides are held in registers for the whole row (one LDS round for the
      // velocities, one for the inverse inertia, one STS round) instead of read-modify-write per component.
      for (int s = 0; s < cnt; s++) {
        int o = L.o_cr + s * PGS_CREC;
        float dinv = SMF(o + 10);
        if (dinv == 0.f) continue;
        int refA = f2i_bits(SMF(o + 16)), refB = f2i_bits(SMF(o + 17));
        f3 n(SMF(o), SMF(o + 1), SMF(o + 2));
        bool fA = (refA & 3) == 1, fB = (refB & 3) == 1;
        int oA = L.o_dv + ND + 6 * (refA >> 2), oB = L.o_dv + ND + 6 * (refB >> 2);
        f3 vlA, vaA, vlB, vaB, aA, aB;
        float jv = 0.f;
        if (fA) {
          vlA = f3(SMF(oA), SMF(oA + 1), SMF(oA + 2)); vaA = f3(SMF(oA + 3), SMF(oA + 4), SMF(oA + 5));
          aA = cross(f3(SMF(o + 3), SMF(o + 4), SMF(o + 5)), n);
          jv += dot(n, vlA) + dot(aA, vaA);
        } else if ((refA & 3) == 2) jv += pgs_side_jv<LANES>(S, e, sm, L, refA, f2i_bits(SMF(o + 18)), -1, n, f3());
        if (fB) {
          vlB = f3(SMF(oB), SMF(oB + 1), SMF(oB + 2)); vaB = f3(SMF(oB + 3), SMF(oB + 4), SMF(oB + 5));
          aB = cross(f3(SMF(o + 6), SMF(o + 7), SMF(o + 8)), n);
          jv -= dot(n, vlB) + dot(aB, vaB);
        } else if ((refB & 3) == 2) jv += pgs_side_jv<LANES>(S, e, sm, L, refB, f2i_bits(SMF(o + 19)), -1, -n, f3());
        float lam = SMF(L.o_lam + 3 * s);
        float dl = SMF(o + 9) - jv * dinv;
        float sum = lam + dl;
        if (sum < 0.f) { dl = -lam; sum = 0.f; }
        SMF(L.o_lam + 3 * s) = sum;
        if (fA) {
          int fi = L.o_fi + 6 * (refA >> 2);
          s3 Ii; Ii.xx = SMF(fi); Ii.yy = SMF(fi + 1); Ii.zz = SMF(fi + 2); Ii.xy = SMF(fi + 3); Ii.xz = SMF(fi + 4); Ii.yz = SMF(fi + 5);
          float k = SMF(L.o_fm + (refA >> 2)) * dl;
          f3 ia = mul(Ii, aA);
          SMF(oA) = vlA.x + n.x * k; SMF(oA + 1) = vlA.y + n.y * k; SMF(oA + 2) = vlA.z + n.z * k;
          SMF(oA + 3) = vaA.x + ia.x * dl; SMF(oA + 4) = vaA.y + ia.y * dl; SMF(oA + 5) = vaA.z + ia.z * dl;
        } else if ((refA & 3) == 2) pgs_side_apply<LANES>(S, e, sm, L, refA, f2i_bits(SMF(o + 18)), -1, n, f3(), dl);
        if (fB) {
          int fi = L.o_fi + 6 * (refB >> 2);
          s3 Ii; Ii.xx = SMF(fi); Ii.yy = SMF(fi + 1); Ii.zz = SMF(fi + 2); Ii.xy = SMF(fi + 3); Ii.xz = SMF(fi + 4); Ii.yz = SMF(fi + 5);
          float k = SMF(L.o_fm + (refB >> 2)) * dl;
          f3 ia = mul(Ii, aB);
          SMF(oB) = vlB.x - n.x * k; SMF(oB + 1) = vlB.y - n.y * k; SMF(oB + 2) = vlB.z - n.z * k;
          SMF(oB + 3) = vaB.x - ia.x * dl; SMF(oB + 4) = vaB.y - ia.y * dl; SMF(oB + 5) = vaB.z - ia.z * dl;
        } else if ((refB & 3) == 2) pgs_side_apply<LANES>(S, e, sm, L, refB, f2i_bits(SMF(o + 19)), -1, -n, f3(), dl);
        resid = fmaxf(resid, dl * dl);
      }
      // friction (two directions per contact, cone or pyramid): both directions share one load / store
      // round of the two bodies' velocities
      for (int s = 0; s < cnt; s++) {
        int o = L.o_cr + s * PGS_CREC;
        if (SMF(o + 10) == 0.f) continue;
        float l1 = SMF(L.o_lam + 3 * s + 1), l2 = SMF(L.o_lam + 3 * s + 2);
        float lim = SMF(o + 15) * SMF(L.o_lam + 3 * s);
        if (lim <= 0.f && l1 == 0.f && l2 == 0.f) continue;
        int refA = f2i_bits(SMF(o + 16)), refB = f2i_bits(SMF(o + 17)), asA = f2i_bits(SMF(o + 18)), asB = f2i_bits(SMF(o + 19));
        f3 n(SMF(o), SMF(o + 1), SMF(o + 2));
        f3 t1, t2; plane_space(n, t1, t2);
        bool fA = (refA & 3) == 1, fB = (refB & 3) == 1;
        int oA = L.o_dv + ND + 6 * (refA >> 2), oB = L.o_dv + ND + 6 * (refB >> 2);
        f3 vlA, vaA, vlB, vaB, a1A, a2A, a1B, a2B;
        float jv1 = 0.f, jv2 = 0.f;
        if (fA) {
          vlA = f3(SMF(oA), SMF(oA + 1), SMF(oA + 2)); vaA = f3(SMF(oA + 3), SMF(oA + 4), SMF(oA + 5));
          f3 rA(SMF(o + 3), SMF(o + 4), SMF(o + 5));
          a1A = cross(rA, t1); a2A = cross(rA, t2);
          jv1 += dot(t1, vlA) + dot(a1A, vaA); jv2 += dot(t2, vlA) + dot(a2A, vaA);
        } else if ((refA & 3) == 2) {
          jv1 += pgs_side_jv<LANES>(S, e, sm, L, refA, asA + 1, -1, t1, f3()); jv2 += pgs_side_jv<LANES>(S, e, sm, L, refA, asA + 2, -1, t2, f3());
        }
        if (fB) {
          vlB = f3(SMF(oB), SMF(oB + 1), SMF(oB + 2)); vaB = f3(SMF(oB + 3), SMF(oB + 4), SMF(oB + 5));
          f3 rB(SMF(o + 6), SMF(o + 7), SMF(o + 8));
          a1B = cross(rB, t1); a2B = cross(rB, t2);
          jv1 -= dot(t1, vlB) + dot(a1B, vaB); jv2 -= dot(t2, vlB) + dot(a2B, vaB);
        } else if ((refB & 3) == 2) {
          jv1 += pgs_side_jv<LANES>(S, e, sm, L, refB, asB + 1, -1, -t1, f3()); jv2 += pgs_side_jv<LANES>(S, e, sm, L, refB, asB + 2, -1, -t2, f3());
        }
        // NOTE: t1 and t2 are solved as one block against the same velocities (block Gauss-Seidel over the
        // pair), exactly as the oracle does
        float s1 = l1 + SMF(o + 11) - jv1 * SMF(o + 12);
        float s2 = l2 + SMF(o + 13) - jv2 * SMF(o + 14);
        if (S.cone) {
          float m2 = s1 * s1 + s2 * s2;
          if (m2 > lim * lim) { float k = lim / sqrtf(m2); s1 *= k; s2 *= k; }
        } else { s1 = clampf(s1, -lim, lim); s2 = clampf(s2, -lim, lim); }
        float d1 = s1 - l1, d2 = s2 - l2;
        SMF(L.o_lam + 3 * s + 1) = s1; SMF(L.o_lam + 3 * s + 2) = s2;
        if (fA) {
          int fi = L.o_fi + 6 * (refA >> 2);
          s3 Ii; Ii.xx = SMF(fi); Ii.yy = SMF(fi + 1); Ii.zz = SMF(fi + 2); Ii.xy = SMF(fi + 3); Ii.xz = SMF(fi + 4); Ii.yz = SMF(fi + 5);
          float im = SMF(L.o_fm + (refA >> 2));
          f3 ia = mul(Ii, a1A * d1 + a2A * d2);
          f3 dl_ = (t1 * d1 + t2 * d2) * im;
          SMF(oA) = vlA.x + dl_.x; SMF(oA + 1) = vlA.y + dl_.y; SMF(oA + 2) = vlA.z + dl_.z;
          SMF(oA + 3) = vaA.x + ia.x; SMF(oA + 4) = vaA.y + ia.y; SMF(oA + 5) = vaA.z + ia.z;
        } else if ((refA & 3) == 2) {
          pgs_side_apply<LANES>(S, e, sm, L, refA, asA + 1, -1, t1, f3(), d1); pgs_side_apply<LANES>(S, e, sm, L, refA, asA + 2, -1, t2, f3(), d2);
        }
        if (fB) {
          int fi = L.o_fi + 6 * (refB >> 2);
          s3 Ii; Ii.xx = SMF(fi); Ii.yy = SMF(fi + 1); Ii.zz = SMF(fi + 2); Ii.xy = SMF(fi + 3); Ii.xz = SMF(fi + 4); Ii.yz = SMF(fi + 5);
          float im = SMF(L.o_fm + (refB >> 2));
          f3 ia = mul(Ii, a1B * d1 + a2B * d2);
          f3 dl_ = (t1 * d1 + t2 * d2) * im;
          SMF(oB) = vlB.x - dl_.x; SMF(oB + 1) = vlB.y - dl_.y; SMF(oB + 2) = vlB.z - dl_.z;
          SMF(oB + 3) = vaB.x - ia.x; SMF(oB + 4) = vaB.y - ia.y; SMF(oB + 5) = vaB.z - ia.z;
        } else if ((refB & 3) == 2) {
          pgs_side_apply<LANES>(S, e, sm, L, refB, asB + 1, -1, -t1, f3(), d1); pgs_side_apply<LANES>(S, e, sm, L, refB, asB + 2, -1, -t2, f3(), d2);
        }
        resid = fmaxf(resid, fmaxf(d1 * d1, d2 * d2));
      }
      if (S.resid_thr > 0.f && resid <= S.resid_thr) done = true;
    }
#if defined(__CUDA_ARCH__)
    if (__all_sync(wmask, done)) break;     // the warp leaves the loop when every env has converged
#else
    if (done) break;
#endif
  }
  // ---- write back
  S.iters_used[e] = used;
#if defined(__CUDA_ARCH__)
  S.pgs_cycles[e] = (int)(clock64() - t_begin);
#endif
  for (int i = 0; i < nvel; i++) S.dv[(size_t)i * N + e] = SMF(L.o_dv + i);
  for (int r = 0; r < 3 * ND; r++) S.dr_lam[(size_t)r * N + e] = SMF(L.o_dr + 5 * r);
  for (int r = 0; r < S.ngr; r++) S.gr_data[((size_t)r * 16 + GR_LAM) * N + e] = SMF(L.o_gr + r * GRW + GR_LAM);
  for (int s = 0; s < cnt; s++) {
    cf_st(S.s_data, s, CF_LAM_N, N, e, SMF(L.o_lam + 3 * s));
    cf_st(S.s_data, s, CF_LAM_T1, N, e, SMF(L.o_lam + 3 * s + 1));
    cf_st(S.s_data, s, CF_LAM_T2, N, e, SMF(L.o_lam + 3 * s + 2));
  }
}
#undef SMF

// ------------------------------------------------------------------ K8: apply deltas, integrate
AG_HDN inline void integrate_body(int e, const SimDev& S, const KP&) {
  const int N = S.N;
  const float dt = S.dt, vmax = S.vmax;
  for (int f = 0; f < S.nf; f++) {
    int b = AG_LDG(S.free_body + f);
    if (S.body_mode[(size_t)b * N + e] != 1) continue;
    int l0 = AG_LDG(S.body_link0 + b);
    int o = S.ND + 6 * f;
    f3 v = ld3(S.base_lin, b, N, e), w = ld3(S.base_ang, b, N, e);
    v = f3(clampf(v.x + ld1(S.dv, o, N, e), -vmax, vmax), clampf(v.y + ld1(S.dv, o + 1, N, e), -vmax, vmax), clampf(v.z + ld1(S.dv, o + 2, N, e), -vmax, vmax));
    w = f3(clampf(w.x + ld1(S.dv, o + 3, N, e), -vmax, vmax), clampf(w.y + ld1(S.dv, o + 4, N, e), -vmax, vmax), clampf(w.z + ld1(S.dv, o + 5, N, e), -vmax, vmax));
    st3(S.base_lin, b, N, e, v); st3(S.base_ang, b, N, e, w);
    f3 com = ld3(S.fcom, f, N, e) + v * dt;
    q4 qn = qnormalize(qmul(qexp(w * dt), ld4(S.base_quat, b, N, e)));
    st4(S.base_quat, b, N, e, qn);
    st3(S.base_pos, b, N, e, com - qrot(qn, tv3(S.link_com, l0)));
  }
  for (int d = 0; d < S.ND; d++) {
    int k = AG_LDG(S.dl_link + d);
    int b = AG_LDG(S.link_body + k);
    if (S.body_mode[(size_t)b * N + e] != 1) continue;
    float qd = clampf(ld1(S.jqd, k, N, e) + ld1(S.dv, d, N, e), -vmax, vmax);
    float qn = ld1(S.jq, k, N, e) + dt * qd;
    if (S.hard_limit[k]) {                      // Human.enforce_joint_limits: teleport back, zero velocity
      float lo = AG_LDG(S.link_lower + k), hi = AG_LDG(S.link_upper + k);
      if (qn < lo) { qn = lo; qd = 0.f; } else if (qn > hi) { qn = hi; qd = 0.f; }
    }
    st1(S.jqd, k, N, e, qd);
    st1(S.jq, k, N, e, qn);
    st1(S.motor_applied, k, N, e, ld1(S.dr_lam, 2 * S.ND + d, N, e) / dt);
  }
  S.c_count[e] = S.c_count[e] > S.maxc ? S.maxc : S.c_count[e];
}
