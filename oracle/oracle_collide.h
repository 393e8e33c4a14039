// oracle_collide.h — convex narrowphase of the CPU oracle: GJK closest points between two vertex
// sets ("cores"), face-plane penetration fallback, sphere-swept radii.
// TEST INFRASTRUCTURE ONLY.
//
// Restates the role of Bullet's btGjkPairDetector + btGjkEpaPenetrationDepthSolver
// (Bullet is not available in this container, so the algorithm below is the
// published GJK distance algorithm with Ericson-style Voronoi simplex reduction, not a transcription).
#pragma once
#include "oracle_math.h"

struct ClosestResult {
  bool overlap;   // cores intersect (then pa/pb/normal/dist come from the penetration fallback)
  V3 pa, pb;      // closest points on core A / core B (world)
  V3 normal;      // unit, from B towards A
  real dist;      // core distance (negative when cores overlap)
  int iters;
};

namespace gjk_detail {

// closest point to the origin on segment ab; returns barycentric (u on a, v on b)
static inline void seg_origin(V3 a, V3 b, real& u, real& v) {
  V3 ab = b - a;
  real t = -dot(a, ab);
  real den = dot(ab, ab);
  if (t <= 0 || den <= 0) { u = 1; v = 0; return; }
  if (t >= den) { u = 0; v = 1; return; }
  v = t / den; u = 1 - v;
}

// closest point to the origin on triangle abc (Ericson, Real-Time Collision Detection 5.1.5)
static inline void tri_origin(V3 a, V3 b, V3 c, real& u, real& v, real& w) {
  V3 ab = b - a, ac = c - a, ap = -a;
  real d1 = dot(ab, ap), d2 = dot(ac, ap);
  if (d1 <= 0 && d2 <= 0) { u = 1; v = 0; w = 0; return; }
  V3 bp = -b;
  real d3 = dot(ab, bp), d4 = dot(ac, bp);
  if (d3 >= 0 && d4 <= d3) { u = 0; v = 1; w = 0; return; }
  real vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { real t = d1 / (d1 - d3); u = 1 - t; v = t; w = 0; return; }
  V3 cp = -c;
  real d5 = dot(ab, cp), d6 = dot(ac, cp);
  if (d6 >= 0 && d5 <= d6) { u = 0; v = 0; w = 1; return; }
  real vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { real t = d2 / (d2 - d6); u = 1 - t; v = 0; w = t; return; }
  real va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) { real t = (d4 - d3) / ((d4 - d3) + (d5 - d6)); u = 0; v = 1 - t; w = t; return; }
  real den = 1 / (va + vb + vc);
  v = vb * den; w = vc * den; u = 1 - v - w;
}

}  // namespace gjk_detail

// GJK distance between convex hulls of vertex sets A and B (world coordinates).
static inline ClosestResult gjk_closest(const V3* A, int nA, const V3* B, int nB) {
  using namespace gjk_detail;
  ClosestResult out;
  out.overlap = false;
  V3 W[4], SA[4], SB[4];
  real lam[4] = {1, 0, 0, 0};
  int n = 0;
  V3 v = A[0] - B[0];
  if (dot(v, v) < real(1e-30)) v = V3(1, 0, 0);
  real scale2 = 0;  // size scale for absolute tolerances
  for (int i = 0; i < nA; i++) scale2 = std::max(scale2, dot(A[i] - A[0], A[i] - A[0]));
  for (int i = 0; i < nB; i++) scale2 = std::max(scale2, dot(B[i] - A[0], B[i] - A[0]));
  const real eps_abs2 = std::max(real(1e-24), scale2 * real(1e-22));
  int it = 0;
  real lower_bound = 0;
  for (; it < 64; it++) {
    int ia = 0, ib = 0;
    real best = -dot(v, A[0]);
    for (int i = 1; i < nA; i++) { real d = -dot(v, A[i]); if (d > best) { best = d; ia = i; } }
    best = dot(v, B[0]);
    for (int i = 1; i < nB; i++) { real d = dot(v, B[i]); if (d > best) { best = d; ib = i; } }
    V3 w = A[ia] - B[ib];
    real vv = dot(v, v);
    real vw = dot(v, w);
    if (n > 0 && vw > 0) lower_bound = std::max(lower_bound, vw / std::sqrt(vv));
    if (n > 0 && vv - vw <= real(1e-12) * vv) break;  // no more progress: v is the closest vector
    bool dup = false;
    for (int i = 0; i < n; i++) if (dot(W[i] - w, W[i] - w) <= eps_abs2) dup = true;
    if (dup) break;
    W[n] = w; SA[n] = A[ia]; SB[n] = B[ib]; n++;
    // closest point of the simplex to the origin, reduce to the supporting sub-simplex
    if (n == 1) { lam[0] = 1; }
    else if (n == 2) {
      real u, t; seg_origin(W[0], W[1], u, t);
      if (t <= 0) { n = 1; lam[0] = 1; }
      else if (u <= 0) { W[0] = W[1]; SA[0] = SA[1]; SB[0] = SB[1]; n = 1; lam[0] = 1; }
      else { lam[0] = u; lam[1] = t; }
    } else if (n == 3) {
      real u, t, s; tri_origin(W[0], W[1], W[2], u, t, s);
      real l3[3] = {u, t, s};
      int m = 0;
      for (int i = 0; i < 3; i++) if (l3[i] > 0) { W[m] = W[i]; SA[m] = SA[i]; SB[m] = SB[i]; lam[m] = l3[i]; m++; }
      n = m;
    } else {
      // tetrahedron: test the four faces; the origin is inside iff it is on the inner side of all
      static const int F[4][4] = {{0, 1, 2, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {1, 3, 2, 0}};
      real bestd = real(1e300);
      int bi = -1; real bl[3] = {0, 0, 0};
      real bestd_all = real(1e300); int bi_all = 0; real bla[3] = {1, 0, 0};
      bool any_outside = false;
      for (int f = 0; f < 4; f++) {
        V3 a = W[F[f][0]], b = W[F[f][1]], c = W[F[f][2]], d = W[F[f][3]];
        V3 nrm = cross(b - a, c - a);
        real sp = -dot(a, nrm);        // origin side
        real sd = dot(d - a, nrm);     // opposite vertex side
        // inside only if clearly on the opposite vertex's side (degenerate tetrahedra count as outside)
        real nl = norm(nrm);
#ifdef ORACLE_FLOAT
        real tol_d = real(1e-5) * nl * norm(d - a), tol_p = real(2e-6) * nl * norm(a);
#else
        real tol_d = real(1e-11) * nl * norm(d - a), tol_p = real(1e-13) * nl * norm(a);
#endif
        bool inside = (sp * sd > 0) && (std::fabs(sd) > tol_d) && (std::fabs(sp) > tol_p);
        real u, t, s; tri_origin(a, b, c, u, t, s);
        V3 p = a * u + b * t + c * s;
        real dd = dot(p, p);
        if (dd < bestd_all) { bestd_all = dd; bi_all = f; bla[0] = u; bla[1] = t; bla[2] = s; }
        if (inside) continue;
        any_outside = true;
        if (dd < bestd) { bestd = dd; bi = f; bl[0] = u; bl[1] = t; bl[2] = s; }
      }
      if (!any_outside) {
        if (lower_bound <= real(1e-9)) { out.overlap = true; break; }   // provably separated otherwise
        bi = bi_all; bl[0] = bla[0]; bl[1] = bla[1]; bl[2] = bla[2];
      }
      V3 tw[3], ta[3], tb[3];
      for (int i = 0; i < 3; i++) { tw[i] = W[F[bi][i]]; ta[i] = SA[F[bi][i]]; tb[i] = SB[F[bi][i]]; }
      int m = 0;
      for (int i = 0; i < 3; i++) if (bl[i] > 0) { W[m] = tw[i]; SA[m] = ta[i]; SB[m] = tb[i]; lam[m] = bl[i]; m++; }
      n = m;
    }
    V3 nv(0, 0, 0);
    for (int i = 0; i < n; i++) nv += W[i] * lam[i];
    v = nv;
    if (dot(v, v) <= eps_abs2) { out.overlap = true; break; }
  }
  out.iters = it;
  if (out.overlap) { out.dist = 0; out.pa = out.pb = V3(); out.normal = V3(0, 0, 1); return out; }
  V3 pa(0, 0, 0), pb(0, 0, 0);
  for (int i = 0; i < n; i++) { pa += SA[i] * lam[i]; pb += SB[i] * lam[i]; }
  out.pa = pa; out.pb = pb;
  V3 d = pa - pb;
  real dn = norm(d);
  out.dist = dn;
  out.normal = dn > 0 ? d * (1 / dn) : V3(0, 0, 1);
  return out;
}

// Penetration fallback when the cores overlap: axis of least penetration over the face normals of
// both cores (planes are (n, d) with n.x <= d inside, world coordinates).  Exact for face contacts
// and for a point core inside a hull; an upper bound on the depth otherwise (edge-edge axes are
// not searched).
static inline void penetration_faces(const V3* A, int nA, const real* PA, int npA,
                                     const V3* B, int nB, const real* PB, int npB, ClosestResult& out) {
  real best = real(-1e300);
  bool found = false;
  for (int k = 0; k < npA; k++) {   // axes out of A
    V3 n(PA[4 * k], PA[4 * k + 1], PA[4 * k + 2]);
    real d = PA[4 * k + 3];
    int jb = 0; real mn = dot(n, B[0]);
    for (int j = 1; j < nB; j++) { real t = dot(n, B[j]); if (t < mn) { mn = t; jb = j; } }
    real sep = mn - d;
    if (sep > best) {
      best = sep; found = true;
      out.normal = -n;               // from B towards A
      out.pb = B[jb];
      out.pa = B[jb] - n * sep;      // projection of the deepest B vertex onto A's face
    }
  }
  for (int k = 0; k < npB; k++) {   // axes out of B
    V3 n(PB[4 * k], PB[4 * k + 1], PB[4 * k + 2]);
    real d = PB[4 * k + 3];
    int ja = 0; real mn = dot(n, A[0]);
    for (int j = 1; j < nA; j++) { real t = dot(n, A[j]); if (t < mn) { mn = t; ja = j; } }
    real sep = mn - d;
    if (sep > best) {
      best = sep; found = true;
      out.normal = n;
      out.pa = A[ja];
      out.pb = A[ja] - n * sep;
    }
  }
  if (!found) { out.normal = V3(0, 0, 1); out.pa = A[0]; out.pb = B[0]; best = 0; }
  out.dist = std::min(best, real(0));
}
