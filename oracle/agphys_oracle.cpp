// agphys_oracle.cpp — CPU restatement of the physics step behind the reference's
// `p.stepSimulation` (assistive_gym/envs/env.py:226) and its read-back calls
// (agents/agent.py:40,52,108,124).
//
// *** TEST INFRASTRUCTURE.  PARITY UNPINNED. ***
// The reference delegates this path to PyBullet (setup.py:21, Zackory/bullet3 fork, unpinned, not
// vendored, not installable here — SURVEY.md §8(c)).  The reference has no tests / golden vectors.
// This file restates the *published algorithms* Bullet uses for the path (Featherstone articulated
// body dynamics, GJK, projected Gauss-Seidel with Bullet's row conventions as recalled in
// SURVEY.md Appendix A).  It is checked against analytic known answers in tests/ (and, for the
// articulated dynamics, against independently derived Lagrange equations), not against PyBullet:
// the PHYSICS is unpinned.  What is pinned to the reference's own code is everything around it --
// tests/golden/make_golden_*.py run the reference's env.step / reset / helper functions on this
// oracle through a pybullet facade and the repo's restatements must reproduce those rollouts
// (DESIGN.md section 5).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.
//
// Deliberately independent of the CUDA implementation:
//   * double precision, array-of-structs, one env at a time,
//   * world-frame spatial algebra about the world origin (CUDA: link-local frames, fp32),
//   * generalized inverse mass via CRBA + Cholesky (CUDA: ABA unit-impulse responses),
//   * fixed joints kept as zero-DoF links (CUDA: merged into their parents).
#include <vector>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <string>
#include "../include/agphys.h"
#include "oracle_math.h"
#include "oracle_collide.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const int MAXD = 48;  // max DoFs of one articulated body

struct Scene {
  int nb, nl, nc, nv, np, npair, ncon;
  std::vector<int> body_link0, body_nlinks;
  std::vector<V3> body_gravity;
  std::vector<int> link_body, link_parent, link_jtype, link_haslimit;
  std::vector<V3> link_axis, link_jpos, link_com, link_inertia;
  std::vector<Quat> link_jquat, link_iquat;
  std::vector<real> link_mass, link_lower, link_upper, link_damping, link_friction;
  std::vector<int> col_link, col_type, col_v0, col_nv, col_p0, col_np;
  std::vector<real> col_radius, col_thresh;
  std::vector<V3> col_center, col_half;     // local bounding box of each collider's core (link frame)
  real max_thresh;
  std::vector<V3> verts;
  std::vector<real> planes;  // 4 per plane
  std::vector<int> pair_link;
  std::vector<int> con_link;
  std::vector<V3> con_pivot;
  std::vector<Quat> con_quat;
  std::vector<real> con_maxforce;
  // derived
  std::vector<int> link_col0, link_ncol;   // colliders are grouped by link
  std::vector<real> subtree_mass;
  std::vector<real> link_thresh;           // max col_thresh over the link's colliders
  std::vector<int> link_live;              // joint has a live DoF (template level)
  std::vector<int> link_dof;               // index of the DoF within its body (-1)
  std::vector<int> body_ndof;
  std::vector<int> body_free;              // body is a free rigid body (single live base)
};

struct Contact {
  int col_a, col_b, link_a, link_b;
  V3 pos_a, pos_b, normal;
  real dist;
  real lambda_n, lambda_t1, lambda_t2;
};

struct Env {
  std::vector<V3> base_pos, base_lin, base_ang;   // base LINK frame position, COM linear velocity, angular velocity (world)
  std::vector<Quat> base_quat;
  std::vector<real> q, qd;                        // per link
  std::vector<int> motor_mode, hard_limit;
  std::vector<real> motor_target, motor_kp, motor_kd, motor_maxf, motor_applied, motor_fscale;
  std::vector<real> friction;                     // per link
  std::vector<int> body_mode;                     // 0 inactive, 1 normal, 2 frozen
  // derived
  std::vector<V3> lpos;                           // link frame world pose
  std::vector<Quat> lquat;
  std::vector<V3> wverts;                         // collider vertices in world
  std::vector<real> wplanes;
  std::vector<V3> cmin, cmax, lmin, lmax;         // collider / link AABBs (including radius)
  std::vector<Contact> contacts;
  int overflow;
  int last_iters;
  // warm-start cache: impulses of the last solve, keyed by what the row is
  std::vector<std::pair<unsigned, real>> ws_contacts;   // (collider pair * 4 + manifold point index, normal impulse), ascending key
  std::vector<real> ws_limit, ws_motor, ws_fixed;       // [2 nl] (lower, upper), [nl], [6 ncon]
};

struct Row {
  // two sides; each side addresses a contiguous slice of the env velocity vector
  int off[2], n[2];
  real J[2][MAXD], MiJ[2][MAXD];
  real diag_inv, rhs, lo, hi, lambda;
  int friction_of;     // index of the normal row this friction row depends on (-1)
  real mu;
};

#include "oracle_cloth.h"

struct Sim {
  Scene sc;
  AgConfig cfg;
  int N;
  std::vector<Env> envs;
  bool has_cloth = false;
  OCloth cloth;
  std::vector<OClothEnv> cloth_envs;
};

thread_local std::string g_err;

// ---------------------------------------------------------------------------- scene ingest
V3 rd3(const double* p, int i) { return V3((real)p[3 * i], (real)p[3 * i + 1], (real)p[3 * i + 2]); }
Quat rd4(const double* p, int i) { return Quat((real)p[4 * i], (real)p[4 * i + 1], (real)p[4 * i + 2], (real)p[4 * i + 3]); }

void ingest(Scene& s, const AgSceneDesc* d) {
  s.nb = d->n_bodies; s.nl = d->n_links; s.nc = d->n_colliders; s.nv = d->n_verts; s.np = d->n_planes;
  s.npair = d->n_pairs; s.ncon = d->n_constraints;
  s.body_link0.assign(d->body_link0, d->body_link0 + s.nb);
  s.body_nlinks.assign(d->body_nlinks, d->body_nlinks + s.nb);
  for (int i = 0; i < s.nb; i++) s.body_gravity.push_back(rd3(d->body_gravity, i));
  s.link_body.assign(d->link_body, d->link_body + s.nl);
  s.link_parent.assign(d->link_parent, d->link_parent + s.nl);
  s.link_jtype.assign(d->link_jtype, d->link_jtype + s.nl);
  s.link_haslimit.assign(d->link_haslimit, d->link_haslimit + s.nl);
  for (int i = 0; i < s.nl; i++) {
    s.link_axis.push_back(rd3(d->link_axis, i)); s.link_jpos.push_back(rd3(d->link_jpos, i));
    s.link_com.push_back(rd3(d->link_com, i)); s.link_inertia.push_back(rd3(d->link_inertia, i));
    s.link_jquat.push_back(rd4(d->link_jquat, i)); s.link_iquat.push_back(rd4(d->link_iquat, i));
    s.link_mass.push_back((real)d->link_mass[i]); s.link_lower.push_back((real)d->link_lower[i]);
    s.link_upper.push_back((real)d->link_upper[i]); s.link_damping.push_back((real)d->link_damping[i]);
    s.link_friction.push_back((real)d->link_friction[i]);
  }
  s.col_link.assign(d->col_link, d->col_link + s.nc); s.col_type.assign(d->col_type, d->col_type + s.nc);
  s.col_v0.assign(d->col_v0, d->col_v0 + s.nc); s.col_nv.assign(d->col_nv, d->col_nv + s.nc);
  s.col_p0.assign(d->col_p0, d->col_p0 + s.nc); s.col_np.assign(d->col_np, d->col_np + s.nc);
  s.max_thresh = 0;
  for (int i = 0; i < s.nc; i++) { s.col_center.push_back(rd3(d->col_center, i)); s.col_half.push_back(rd3(d->col_half, i)); s.col_radius.push_back((real)d->col_radius[i]); s.col_thresh.push_back((real)d->col_thresh[i]); s.max_thresh = std::max(s.max_thresh, (real)d->col_thresh[i]); }
  for (int i = 0; i < s.nv; i++) s.verts.push_back(rd3(d->verts, i));
  for (int i = 0; i < 4 * s.np; i++) s.planes.push_back((real)d->planes[i]);
  s.pair_link.assign(d->pair_link, d->pair_link + 2 * s.npair);
  s.con_link.assign(d->con_link, d->con_link + 2 * s.ncon);
  for (int i = 0; i < 2 * s.ncon; i++) { s.con_pivot.push_back(rd3(d->con_pivot, i)); s.con_quat.push_back(rd4(d->con_quat, i)); }
  for (int i = 0; i < s.ncon; i++) s.con_maxforce.push_back((real)d->con_maxforce[i]);
  // derived
  s.link_col0.assign(s.nl, 0); s.link_ncol.assign(s.nl, 0);
  for (int c = s.nc - 1; c >= 0; c--) { s.link_col0[s.col_link[c]] = c; s.link_ncol[s.col_link[c]]++; }
  s.link_thresh.assign(s.nl, 0);
  for (int c = 0; c < s.nc; c++) s.link_thresh[s.col_link[c]] = std::max(s.link_thresh[s.col_link[c]], s.col_thresh[c]);
  s.subtree_mass.assign(s.nl, 0);
  for (int k = s.nl - 1; k >= 0; k--) {
    s.subtree_mass[k] += s.link_mass[k];
    if (s.link_parent[k] >= 0) s.subtree_mass[s.link_parent[k]] += s.subtree_mass[k];
  }
  s.link_live.assign(s.nl, 0); s.link_dof.assign(s.nl, -1); s.body_ndof.assign(s.nb, 0); s.body_free.assign(s.nb, 0);
  for (int b = 0; b < s.nb; b++) {
    int l0 = s.body_link0[b];
    if (s.link_jtype[l0] == AG_JOINT_FREE_BASE && s.link_mass[l0] > 0) s.body_free[b] = 1;
    for (int k = l0 + 1; k < l0 + s.body_nlinks[b]; k++) {
      int jt = s.link_jtype[k];
      if ((jt == AG_JOINT_REVOLUTE || jt == AG_JOINT_PRISMATIC) && s.subtree_mass[k] > 0) {
        s.link_live[k] = 1; s.link_dof[k] = s.body_ndof[b]++;
      }
    }
  }
}

// ---------------------------------------------------------------------------- kinematics
void forward_kinematics(const Scene& s, Env& e) {
  for (int b = 0; b < s.nb; b++) {
    int l0 = s.body_link0[b];
    e.lpos[l0] = e.base_pos[b]; e.lquat[l0] = e.base_quat[b];
    for (int k = l0 + 1; k < l0 + s.body_nlinks[b]; k++) {
      int p = s.link_parent[k];
      V3 jp = e.lpos[p] + qrot(e.lquat[p], s.link_jpos[k]);
      Quat jq = qmul(e.lquat[p], s.link_jquat[k]);
      int jt = s.link_jtype[k];
      if (jt == AG_JOINT_REVOLUTE) jq = qmul(jq, qaxis(s.link_axis[k], e.q[k]));
      else if (jt == AG_JOINT_PRISMATIC) jp = jp + qrot(jq, s.link_axis[k] * e.q[k]);
      e.lpos[k] = jp; e.lquat[k] = qnormalize(jq);
    }
  }
}

M3 world_inertia(const Scene& s, const Env& e, int k) {
  M3 R = qmat(qmul(e.lquat[k], s.link_iquat[k]));
  return R * M3::diag(s.link_inertia[k]) * transpose(R);
}
V3 world_com(const Scene& s, const Env& e, int k) { return e.lpos[k] + qrot(e.lquat[k], s.link_com[k]); }

// spatial velocity (world, about origin) of every link of a fixed-base articulated body
void link_velocities(const Scene& s, const Env& e, int b, std::vector<SV>& v) {
  int l0 = s.body_link0[b];
  v[l0] = SV();
  for (int k = l0 + 1; k < l0 + s.body_nlinks[b]; k++) {
    int p = s.link_parent[k];
    v[k] = v[p];
    if (s.link_live[k]) {
      V3 a = qrot(e.lquat[k], s.link_axis[k]);
      SV S = (s.link_jtype[k] == AG_JOINT_REVOLUTE) ? SV(a, cross(e.lpos[k], a)) : SV(V3(), a);
      v[k] = v[k] + S * e.qd[k];
    }
  }
}

// ---------------------------------------------------------------------------- collision
void update_colliders(const Scene& s, Env& e) {
  for (int c = 0; c < s.nc; c++) {
    int k = s.col_link[c];
    V3 mn(1e30, 1e30, 1e30), mx(-1e30, -1e30, -1e30);
    for (int i = 0; i < s.col_nv[c]; i++) {
      V3 w = e.lpos[k] + qrot(e.lquat[k], s.verts[s.col_v0[c] + i]);
      e.wverts[s.col_v0[c] + i] = w;
      for (int a = 0; a < 3; a++) { mn[a] = std::min(mn[a], w[a]); mx[a] = std::max(mx[a], w[a]); }
    }
    for (int i = 0; i < s.col_np[c]; i++) {
      int p = s.col_p0[c] + i;
      V3 n = qrot(e.lquat[k], V3(s.planes[4 * p], s.planes[4 * p + 1], s.planes[4 * p + 2]));
      e.wplanes[4 * p] = n.x; e.wplanes[4 * p + 1] = n.y; e.wplanes[4 * p + 2] = n.z;
      e.wplanes[4 * p + 3] = s.planes[4 * p + 3] + dot(n, e.lpos[k]);
    }
    real r = s.col_radius[c];
    if (s.col_type[c] == AG_COL_HALFSPACE) {
      mn = V3(-1e30, -1e30, -1e30); mx = V3(1e30, 1e30, 1e30);
      const real* P = &e.wplanes[4 * s.col_p0[c]];     // axis-aligned half-spaces are bounded on one side
      for (int a = 0; a < 3; a++) { if (P[a] > real(0.999999)) mx[a] = P[3]; else if (P[a] < real(-0.999999)) mn[a] = -P[3]; }
    }
    e.cmin[c] = mn - V3(r, r, r); e.cmax[c] = mx + V3(r, r, r);
  }
  for (int k = 0; k < s.nl; k++) {
    V3 mn(1e30, 1e30, 1e30), mx(-1e30, -1e30, -1e30);
    for (int c = s.link_col0[k]; c < s.link_col0[k] + s.link_ncol[k]; c++)
      for (int a = 0; a < 3; a++) { mn[a] = std::min(mn[a], e.cmin[c][a]); mx[a] = std::max(mx[a], e.cmax[c][a]); }
    e.lmin[k] = mn; e.lmax[k] = mx;
  }
}

inline bool aabb_overlap(V3 amin, V3 amax, V3 bmin, V3 bmax, real margin) {
  for (int a = 0; a < 3; a++) if (amin[a] > bmax[a] + margin || bmin[a] > amax[a] + margin) return false;
  return true;
}

// One-shot manifold: besides the primary closest-point pair, vertices of one core that lie over the
// other core's supporting face (within `tol` of the primary distance) become extra contact points, so
// that face-face / edge-face resting contacts get up to 4 points in a single frame.  Plays the role of
// Bullet's persistent 4-point manifold without per-pair state.
struct Cand { V3 pa, pb, n; real d; };

void face_candidates(const V3* V, int nV, real rV, const real* P, int nP, real rP, V3 n_to_v, real d_primary, real tol, real max_dist,
                     bool v_is_a, std::vector<Cand>& out) {
  // supporting face of the plane owner along n_to_v (direction from the face owner towards V's owner)
  int kf = -1; real best = real(0.98);
  for (int k = 0; k < nP; k++) { real al = P[4 * k] * n_to_v.x + P[4 * k + 1] * n_to_v.y + P[4 * k + 2] * n_to_v.z; if (al > best) { best = al; kf = k; } }
  if (kf < 0) return;
  V3 nf(P[4 * kf], P[4 * kf + 1], P[4 * kf + 2]); real df = P[4 * kf + 3];
  for (int i = 0; i < nV; i++) {
    real h = dot(nf, V[i]) - df;             // core height of the vertex above the face
    real d = h - rV - rP;
    if (d > d_primary + tol || d > max_dist) continue;
    V3 proj = V[i] - nf * h;
    bool inside = true;
    for (int k = 0; k < nP && inside; k++) if (k != kf && P[4 * k] * proj.x + P[4 * k + 1] * proj.y + P[4 * k + 2] * proj.z - P[4 * k + 3] > real(1e-6)) inside = false;
    if (!inside) continue;
    Cand c; c.d = d;
    V3 on_v = V[i] - nf * rV, on_f = proj + nf * rP;
    if (v_is_a) { c.pa = on_v; c.pb = on_f; c.n = nf; } else { c.pa = on_f; c.pb = on_v; c.n = -nf; }
    if (out.size() >= 13) {          // pool of 12 (+ the primary at index 0): replace the shallowest if deeper
      size_t wi = 1; for (size_t q = 2; q < out.size(); q++) if (out[q].d > out[wi].d) wi = q;
      if (d >= out[wi].d) continue;
      out[wi] = c;
    } else out.push_back(c);
  }
}

// contacts between two colliders: appends up to 4 points with surface distance <= max_dist
int collide_pair(const Scene& s, const Env& e, int ca, int cb, real max_dist, bool manifold, Contact* outc) {
  int ta = s.col_type[ca], tb = s.col_type[cb];
  real ra = s.col_radius[ca], rb = s.col_radius[cb];
  const V3* A = &e.wverts[s.col_v0[ca]]; const V3* B = &e.wverts[s.col_v0[cb]];
  int nA = s.col_nv[ca], nB = s.col_nv[cb];
  const real* PA = &e.wplanes[4 * s.col_p0[ca]]; const real* PB = &e.wplanes[4 * s.col_p0[cb]];
  int npA = s.col_np[ca], npB = s.col_np[cb];
  std::vector<Cand> cand;
  Cand pr;
  if (ta == AG_COL_HALFSPACE || tb == AG_COL_HALFSPACE) {
    if (ta == tb) return 0;
    bool flip = (ta == AG_COL_HALFSPACE);        // the half-space is A
    const V3* V = flip ? B : A; int nV = flip ? nB : nA; real rv = flip ? rb : ra;
    const real* P = flip ? PA : PB;
    V3 pn(P[0], P[1], P[2]);
    int j = 0; real mn = dot(pn, V[0]);
    for (int i = 1; i < nV; i++) { real t = dot(pn, V[i]); if (t < mn) { mn = t; j = i; } }
    pr.d = mn - P[3] - rv;
    if (pr.d > max_dist) return 0;
    V3 on_shape = V[j] - pn * rv, on_plane = V[j] - pn * (mn - P[3]);
    if (!flip) { pr.pa = on_shape; pr.pb = on_plane; pr.n = pn; } else { pr.pa = on_plane; pr.pb = on_shape; pr.n = -pn; }
    cand.push_back(pr);
    if (manifold && nV > 1) face_candidates(V, nV, rv, P, 1, 0, pn, pr.d, max_dist * real(0.5), max_dist, !flip, cand);
  } else {
    ClosestResult r = gjk_closest(A, nA, B, nB);
    if (r.overlap) penetration_faces(A, nA, PA, npA, B, nB, PB, npB, r);
    pr.d = r.dist - ra - rb;
    if (pr.d > max_dist) return 0;
    pr.n = r.normal; pr.pa = r.pa - r.normal * ra; pr.pb = r.pb + r.normal * rb;
    cand.push_back(pr);
    if (manifold) {
      if (npB > 0 && nA > 1) face_candidates(A, nA, ra, PB, npB, rb, r.normal, pr.d, max_dist * real(0.5), max_dist, true, cand);
      if (npA > 0 && nB > 1) face_candidates(B, nB, rb, PA, npA, ra, -r.normal, pr.d, max_dist * real(0.5), max_dist, false, cand);
    }
  }
  // Manifold selection (same rule as the CUDA side): the GJK primary is arbitrary within a flat
  // contact patch, so when feature candidates exist only they are used: deepest first, then greedily
  // the candidate farthest from the chosen set.  cand[0] is the primary, cand[1..] the pool.
  int chosen[4]; int nc = 0;
  std::vector<char> used(cand.size(), 0); used[0] = 1;
  if (cand.size() == 1) { chosen[nc++] = 0; }
  else {
    size_t first = 1;
    for (size_t i = 2; i < cand.size(); i++) if (cand[i].d < cand[first].d) first = i;
    chosen[nc++] = (int)first; used[first] = 1;
    while (nc < 4) {
      int bi = -1; real bd = real(1e-8);
      for (size_t i = 1; i < cand.size(); i++) {
        if (used[i]) continue;
        real md = real(1e30);
        for (int k = 0; k < nc; k++) { V3 dd = cand[i].pa - cand[chosen[k]].pa; md = std::min(md, dot(dd, dd)); }
        if (md > bd) { bd = md; bi = (int)i; }
      }
      if (bi < 0) break;
      used[bi] = 1; chosen[nc++] = bi;
    }
  }
  for (int k = 0; k < nc; k++) {
    const Cand& c = cand[chosen[k]];
    Contact& o = outc[k];
    o.col_a = ca; o.col_b = cb; o.link_a = s.col_link[ca]; o.link_b = s.col_link[cb];
    o.pos_a = c.pa; o.pos_b = c.pb; o.normal = c.n; o.dist = c.d;
    o.lambda_n = o.lambda_t1 = o.lambda_t2 = 0;
  }
  return nc;
}

void detect_contacts(const Scene& s, const AgConfig& cfg, Env& e) {
  e.contacts.clear();
  real fac = (real)cfg.contact_threshold;
  for (int p = 0; p < s.npair; p++) {
    int la = s.pair_link[2 * p], lb = s.pair_link[2 * p + 1];
    if (e.body_mode[s.link_body[la]] == 0 || e.body_mode[s.link_body[lb]] == 0) continue;
    if (!aabb_overlap(e.lmin[la], e.lmax[la], e.lmin[lb], e.lmax[lb], fac * std::min(s.link_thresh[la], s.link_thresh[lb]))) continue;
    for (int ca = s.link_col0[la]; ca < s.link_col0[la] + s.link_ncol[la]; ca++) {
      if (!aabb_overlap(e.cmin[ca], e.cmax[ca], e.lmin[lb], e.lmax[lb], fac * std::min(s.col_thresh[ca], s.link_thresh[lb]))) continue;
      for (int cb = s.link_col0[lb]; cb < s.link_col0[lb] + s.link_ncol[lb]; cb++) {
        real thr = fac * std::min(s.col_thresh[ca], s.col_thresh[cb]);   // size-relative breaking threshold
        if (!aabb_overlap(e.cmin[ca], e.cmax[ca], e.cmin[cb], e.cmax[cb], thr)) continue;
        Contact c[4];
        int n = collide_pair(s, e, ca, cb, thr, true, c);
        for (int i = 0; i < n; i++) e.contacts.push_back(c[i]);
      }
    }
  }
}

// ---------------------------------------------------------------------------- dynamics
struct BodyDyn {              // per articulated body scratch
  int ndof;
  real M[MAXD][MAXD];         // joint-space inertia, then its Cholesky factor (lower)
};

bool body_moves(const Scene& s, const Env& e, int b) {
  return e.body_mode[b] == 1 && (s.body_free[b] || s.body_ndof[b] > 0);
}

SV joint_axis(const Scene& s, const Env& e, int k) {
  V3 a = qrot(e.lquat[k], s.link_axis[k]);
  return (s.link_jtype[k] == AG_JOINT_REVOLUTE) ? SV(a, cross(e.lpos[k], a)) : SV(V3(), a);
}

// Featherstone ABA for a fixed-base tree; adds dt*qdd to qd.
void aba_fixed_base(const Scene& s, const AgConfig& cfg, Env& e, int b, real dt) {
  int l0 = s.body_link0[b], nl = s.body_nlinks[b];
  std::vector<SV> v(nl), c(nl), pA(nl), U(nl), S(nl), a(nl);
  std::vector<SI> IA(nl);
  std::vector<real> D(nl), u(nl);
  V3 g = s.body_gravity[b];
  real kl = (real)cfg.linear_damping, ka = (real)cfg.angular_damping;
  v[0] = SV(); c[0] = SV();
  IA[0] = SI(); pA[0] = SV();
  for (int i = 1; i < nl; i++) {
    int k = l0 + i, p = s.link_parent[k] - l0;
    v[i] = v[p]; c[i] = SV();
    if (s.link_live[k]) {
      S[i] = joint_axis(s, e, k);
      SV vj = S[i] * e.qd[k];
      v[i] = v[i] + vj;
      c[i] = crm(v[i], vj);
    }
    real m = s.link_mass[k];
    V3 com = world_com(s, e, k);
    M3 Ic = world_inertia(s, e, k);
    IA[i] = rigid_inertia(m, com, Ic);
    pA[i] = crf(v[i], IA[i] * v[i]);
    // Bullet-style velocity damping applied at the COM
    V3 vc = v[i].l + cross(v[i].a, com);
    V3 w = v[i].a;
    V3 f = vc * (-m * (kl + kl * norm(vc)));
    V3 n = (Ic * w) * (-(ka + ka * norm(w)));
    pA[i] = pA[i] - SV(n + cross(com, f), f);
  }
  for (int i = nl - 1; i >= 1; i--) {
    int k = l0 + i, p = s.link_parent[k] - l0;
    if (s.link_live[k]) {
      U[i] = IA[i] * S[i];
      D[i] = sdot(S[i], U[i]);
      real tau = -s.link_damping[k] * e.qd[k];
      u[i] = tau - sdot(S[i], pA[i]);
      SI Ia = sub_outer(IA[i], U[i], 1 / D[i]);
      SV pa = pA[i] + Ia * c[i] + U[i] * (u[i] / D[i]);
      IA[p] = IA[p] + Ia; pA[p] = pA[p] + pa;
    } else {
      IA[p] = IA[p] + IA[i]; pA[p] = pA[p] + pA[i] + IA[i] * c[i];
    }
  }
  a[0] = SV(V3(), -g);
  for (int i = 1; i < nl; i++) {
    int k = l0 + i, p = s.link_parent[k] - l0;
    SV ap = a[p] + c[i];
    if (s.link_live[k]) {
      real qdd = (u[i] - sdot(U[i], ap)) / D[i];
      a[i] = ap + S[i] * qdd;
      e.qd[k] += dt * qdd;
    } else a[i] = ap;
  }
}

// CRBA joint-space inertia + Cholesky for a fixed-base tree
void crba_factor(const Scene& s, const Env& e, int b, BodyDyn& bd) {
  int l0 = s.body_link0[b], nl = s.body_nlinks[b];
  int nd = s.body_ndof[b];
  bd.ndof = nd;
  std::vector<SI> Ic(nl);
  std::vector<SV> S(nl);
  for (int i = 1; i < nl; i++) {
    int k = l0 + i;
    Ic[i] = rigid_inertia(s.link_mass[k], world_com(s, e, k), world_inertia(s, e, k));
    if (s.link_live[k]) S[i] = joint_axis(s, e, k);
  }
  for (int i = 0; i < nd; i++) for (int j = 0; j < nd; j++) bd.M[i][j] = 0;
  for (int i = nl - 1; i >= 1; i--) {
    int k = l0 + i, p = s.link_parent[k] - l0;
    if (p >= 1) Ic[p] = Ic[p] + Ic[i];
    if (!s.link_live[k]) continue;
    SV F = Ic[i] * S[i];
    int di = s.link_dof[k];
    bd.M[di][di] = sdot(S[i], F);
    int j = p;
    while (j >= 1) {
      int kj = l0 + j;
      if (s.link_live[kj]) { int dj = s.link_dof[kj]; bd.M[di][dj] = bd.M[dj][di] = sdot(S[j], F); }
      j = s.link_parent[kj] - l0;
    }
  }
  // Cholesky M = L L^T in place (lower)
  for (int j = 0; j < nd; j++) {
    real d = bd.M[j][j];
    for (int k = 0; k < j; k++) d -= bd.M[j][k] * bd.M[j][k];
    d = std::sqrt(d);
    bd.M[j][j] = d;
    for (int i = j + 1; i < nd; i++) {
      real t = bd.M[i][j];
      for (int k = 0; k < j; k++) t -= bd.M[i][k] * bd.M[j][k];
      bd.M[i][j] = t / d;
    }
  }
}
void chol_solve(const BodyDyn& bd, const real* rhs, real* x) {
  int n = bd.ndof;
  real y[MAXD];
  for (int i = 0; i < n; i++) { real t = rhs[i]; for (int k = 0; k < i; k++) t -= bd.M[i][k] * y[k]; y[i] = t / bd.M[i][i]; }
  for (int i = n - 1; i >= 0; i--) { real t = y[i]; for (int k = i + 1; k < n; k++) t -= bd.M[k][i] * x[k]; x[i] = t / bd.M[i][i]; }
}

// ---------------------------------------------------------------------------- constraint rows
struct Solver {
  const Scene& s; const AgConfig& cfg; Env& e;
  std::vector<int> body_off;          // offset of each body in the velocity vector (-1: immovable)
  std::vector<BodyDyn> dyn;
  std::vector<real> vel, dv;          // current velocities (after free update), solver deltas
  std::vector<M3> free_Iinv;          // world inverse inertia of free bodies
  std::vector<V3> free_com;
  std::vector<Row> rows;
  Solver(const Scene& s_, const AgConfig& c_, Env& e_) : s(s_), cfg(c_), e(e_) {}

  // fill one side of a row for a unit (force `lin` at world point p) + (torque `ang`) acting on link k
  void side(Row& r, int sd, int k, V3 p, V3 lin, V3 ang) {
    int b = s.link_body[k];
    r.off[sd] = 0; r.n[sd] = 0;
    if (body_off[b] < 0) return;
    if (s.body_free[b]) {
      r.off[sd] = body_off[b]; r.n[sd] = 6;
      V3 t = cross(p - free_com[b], lin) + ang;
      real* J = r.J[sd]; real* Mi = r.MiJ[sd];
      J[0] = lin.x; J[1] = lin.y; J[2] = lin.z; J[3] = t.x; J[4] = t.y; J[5] = t.z;
      real im = 1 / s.link_mass[s.body_link0[b]];
      V3 it = free_Iinv[b] * t;
      Mi[0] = lin.x * im; Mi[1] = lin.y * im; Mi[2] = lin.z * im; Mi[3] = it.x; Mi[4] = it.y; Mi[5] = it.z;
    } else {
      int nd = s.body_ndof[b];
      r.off[sd] = body_off[b]; r.n[sd] = nd;
      real* J = r.J[sd];
      for (int i = 0; i < nd; i++) J[i] = 0;
      int j = k;
      while (j >= 0) {
        if (s.link_live[j]) {
          V3 a = qrot(e.lquat[j], s.link_axis[j]);
          if (s.link_jtype[j] == AG_JOINT_REVOLUTE) J[s.link_dof[j]] = dot(lin, cross(a, p - e.lpos[j])) + dot(ang, a);
          else J[s.link_dof[j]] = dot(lin, a);
        }
        j = s.link_parent[j];
      }
      chol_solve(dyn[b], J, r.MiJ[sd]);
    }
  }
  real jv(const Row& r, const std::vector<real>& v) const {
    real t = 0;
    for (int sd = 0; sd < 2; sd++) for (int i = 0; i < r.n[sd]; i++) t += r.J[sd][i] * v[r.off[sd] + i];
    return t;
  }
  bool finish(Row& r) {
    // two links of the same articulated body (self-collision): both sides act on the same dof vector, so the row is
    // ONE side with J = J_a + J_b (the diagonal needs the cross terms J_a M^-1 J_b^T)
    if (r.n[0] > 0 && r.n[1] == r.n[0] && r.off[0] == r.off[1]) {
      for (int i = 0; i < r.n[0]; i++) { r.J[0][i] += r.J[1][i]; r.MiJ[0][i] += r.MiJ[1][i]; }
      r.n[1] = 0;
    }
    real d = 0;
    for (int sd = 0; sd < 2; sd++) for (int i = 0; i < r.n[sd]; i++) d += r.J[sd][i] * r.MiJ[sd][i];
    if (d <= real(1e-30)) return false;
    r.diag_inv = 1 / d; r.lambda = 0; r.friction_of = -1; r.mu = 0;
    return true;
  }
};

void plane_space(V3 n, V3& t1, V3& t2) {
  // same construction as the CUDA side is NOT required; any orthonormal pair spans the tangent plane,
  // but a fixed rule keeps GPU/oracle rows comparable: Bullet's btPlaneSpace1.
  const real SQRT12 = real(0.7071067811865475244);
  if (std::fabs(n.z) > SQRT12) {
    real a = n.y * n.y + n.z * n.z; real k = 1 / std::sqrt(a);
    t1 = V3(0, -n.z * k, n.y * k);
    t2 = V3(a * k, -n.x * t1.z, n.x * t1.y);
  } else {
    real a = n.x * n.x + n.y * n.y; real k = 1 / std::sqrt(a);
    t1 = V3(-n.y * k, n.x * k, 0);
    t2 = V3(-n.z * t1.y, n.z * t1.x, a * k);
  }
}

void step_env(const Scene& s, const AgConfig& cfg, Env& e, const OCloth* cloth = nullptr, OClothEnv* cloth_env = nullptr) {
  const real dt = (real)(cfg.dt / std::max(1, cfg.num_substeps));
  for (int sub = 0; sub < std::max(1, cfg.num_substeps); sub++) {
    // 1. kinematics + collision detection at the current positions
    forward_kinematics(s, e);
    update_colliders(s, e);
    detect_contacts(s, cfg, e);
    e.overflow = (int)e.contacts.size() > cfg.max_contacts;
    if (e.overflow) e.contacts.resize(cfg.max_contacts);

    // 2. unconstrained velocity update (gravity, gyroscopic, damping)
    Solver so(s, cfg, e);
    so.body_off.assign(s.nb, -1); so.dyn.resize(s.nb); so.free_Iinv.resize(s.nb); so.free_com.resize(s.nb);
    int nvel = 0;
    real vmax = (real)cfg.max_coord_velocity;
    for (int b = 0; b < s.nb; b++) {
      if (!body_moves(s, e, b)) continue;
      int l0 = s.body_link0[b];
      if (s.body_free[b]) {
        real m = s.link_mass[l0];
        M3 Iw = world_inertia(s, e, l0);
        M3 Iinv = inverse(Iw);
        V3 v = e.base_lin[b], w = e.base_ang[b];
        real kl = (real)cfg.linear_damping, ka = (real)cfg.angular_damping;
        V3 acc = s.body_gravity[b] - v * (kl + kl * norm(v));
        V3 tau = -((Iw * w) * (ka + ka * norm(w)));
        if (cfg.gyroscopic) tau = tau - cross(w, Iw * w);
        v = v + acc * dt; w = w + (Iinv * tau) * dt;
        for (int a = 0; a < 3; a++) { v[a] = std::min(vmax, std::max(-vmax, v[a])); w[a] = std::min(vmax, std::max(-vmax, w[a])); }
        e.base_lin[b] = v; e.base_ang[b] = w;
        so.free_Iinv[b] = Iinv; so.free_com[b] = world_com(s, e, l0);
        so.body_off[b] = nvel; nvel += 6;
      } else {
        aba_fixed_base(s, cfg, e, b, dt);
        for (int k = l0 + 1; k < l0 + s.body_nlinks[b]; k++) if (s.link_live[k]) e.qd[k] = std::min(vmax, std::max(-vmax, e.qd[k]));
        crba_factor(s, e, b, so.dyn[b]);
        so.body_off[b] = nvel; nvel += s.body_ndof[b];
      }
    }
    so.vel.assign(nvel, 0); so.dv.assign(nvel, 0);
    for (int b = 0; b < s.nb; b++) {
      if (so.body_off[b] < 0) continue;
      int o = so.body_off[b], l0 = s.body_link0[b];
      if (s.body_free[b]) { for (int a = 0; a < 3; a++) { so.vel[o + a] = e.base_lin[b][a]; so.vel[o + 3 + a] = e.base_ang[b][a]; } }
      else for (int k = l0 + 1; k < l0 + s.body_nlinks[b]; k++) if (s.link_live[k]) so.vel[o + s.link_dof[k]] = e.qd[k];
    }

    // 3. constraint rows.  Order: joint limits, motors, fixed constraints, contact normals, friction.
    std::vector<Row>& rows = so.rows;
    std::vector<int> motor_row(s.nl, -1), limit_row(2 * s.nl, -1), fixed_row(6 * s.ncon, -1);
    real erp = (real)cfg.erp;
    const real wsc = (real)cfg.warmstart_contact, wsj = (real)cfg.warmstart_joint;
    if ((int)e.ws_limit.size() != 2 * s.nl) { e.ws_limit.assign(2 * s.nl, 0); e.ws_motor.assign(s.nl, 0); e.ws_fixed.assign(6 * s.ncon, 0); }
    for (int k = 0; k < s.nl; k++) {     // joint limits: a row only while the limit is violated
      if (!s.link_live[k] || !s.link_haslimit[k] || so.body_off[s.link_body[k]] < 0) continue;
      for (int sgn = 0; sgn < 2; sgn++) {
        real pen = sgn == 0 ? (e.q[k] - s.link_lower[k]) : (s.link_upper[k] - e.q[k]);
        if (pen > 0) continue;
        Row r; int b = s.link_body[k];
        r.off[0] = so.body_off[b]; r.n[0] = s.body_ndof[b]; r.off[1] = 0; r.n[1] = 0;
        for (int i = 0; i < r.n[0]; i++) r.J[0][i] = 0;
        r.J[0][s.link_dof[k]] = sgn == 0 ? 1 : -1;
        chol_solve(so.dyn[b], r.J[0], r.MiJ[0]);
        if (!so.finish(r)) continue;
        real rel = so.jv(r, so.vel);
        r.rhs = (-pen * erp / dt - rel) * r.diag_inv; r.lo = 0; r.hi = real(1e30);
        r.lambda = std::max(real(0), wsj * e.ws_limit[2 * k + sgn]);
        limit_row[2 * k + sgn] = (int)rows.size();
        rows.push_back(r);
      }
    }
    for (int k = 0; k < s.nl; k++) {     // joint motors
      e.motor_applied[k] = 0;
      if (!s.link_live[k] || e.motor_mode[k] == AG_MOTOR_OFF || so.body_off[s.link_body[k]] < 0) continue;
      real maxi = e.motor_maxf[k] * dt * e.motor_fscale[k];       // Human.strength scales the force limit (human.py:86,126)
      if (maxi <= 0) continue;
      Row r; int b = s.link_body[k];
      r.off[0] = so.body_off[b]; r.n[0] = s.body_ndof[b]; r.off[1] = 0; r.n[1] = 0;
      for (int i = 0; i < r.n[0]; i++) r.J[0][i] = 0;
      r.J[0][s.link_dof[k]] = 1;
      chol_solve(so.dyn[b], r.J[0], r.MiJ[0]);
      if (!so.finish(r)) continue;
      real qd = e.qd[k];
      real vt;
      if (e.motor_mode[k] == AG_MOTOR_POSITION) vt = e.motor_kp[k] * (e.motor_target[k] - e.q[k]) / dt + qd + e.motor_kd[k] * (0 - qd);
      else vt = e.motor_target[k];
      r.rhs = (vt - so.jv(r, so.vel)) * r.diag_inv; r.lo = -maxi; r.hi = maxi;
      r.lambda = std::min(maxi, std::max(-maxi, wsj * e.ws_motor[k]));
      motor_row[k] = (int)rows.size();
      rows.push_back(r);
    }
    for (int c = 0; c < s.ncon; c++) {   // fixed constraints: 3 linear + 3 angular rows, world axes
      int ka = s.con_link[2 * c], kb = s.con_link[2 * c + 1];
      if (e.body_mode[s.link_body[ka]] == 0 || e.body_mode[s.link_body[kb]] == 0) continue;
      V3 pa = e.lpos[ka] + qrot(e.lquat[ka], s.con_pivot[2 * c]);
      V3 pb = e.lpos[kb] + qrot(e.lquat[kb], s.con_pivot[2 * c + 1]);
      Quat fa = qmul(e.lquat[ka], s.con_quat[2 * c]), fb = qmul(e.lquat[kb], s.con_quat[2 * c + 1]);
      Quat qe = qmul(fa, qconj(fb));
      if (qe.w < 0) qe = Quat(-qe.x, -qe.y, -qe.z, -qe.w);
      V3 perr = pa - pb, aerr(2 * qe.x, 2 * qe.y, 2 * qe.z);
      real maxi = s.con_maxforce[c] * dt;
      for (int i = 0; i < 6; i++) {
        V3 ax(0, 0, 0); ax[i % 3] = 1;
        Row r;
        if (i < 3) { so.side(r, 0, ka, pa, ax, V3()); so.side(r, 1, kb, pb, -ax, V3()); }
        else { so.side(r, 0, ka, pa, V3(), ax); so.side(r, 1, kb, pb, V3(), -ax); }
        if (!so.finish(r)) continue;
        real err = i < 3 ? perr[i] : aerr[i - 3];
        r.rhs = (-err * erp / dt - so.jv(r, so.vel)) * r.diag_inv; r.lo = -maxi; r.hi = maxi;
        r.lambda = std::min(maxi, std::max(-maxi, wsj * e.ws_fixed[6 * c + i]));
        fixed_row[6 * c + i] = (int)rows.size();
        rows.push_back(r);
      }
    }
    int first_contact_row = (int)rows.size();
    std::vector<int> crow(e.contacts.size(), -1);
    std::vector<unsigned> ckeys(e.contacts.size());
    for (size_t ci = 0, run = 0; ci < e.contacts.size(); ci++) {     // manifold point index = position within the collider pair's run
      if (ci > 0 && e.contacts[ci].col_a == e.contacts[ci - 1].col_a && e.contacts[ci].col_b == e.contacts[ci - 1].col_b) run++; else run = 0;
      ckeys[ci] = ((unsigned)e.contacts[ci].col_a * (unsigned)s.nc + (unsigned)e.contacts[ci].col_b) * 4u + (unsigned)run;
    }
    for (size_t ci = 0; ci < e.contacts.size(); ci++) {
      Contact& c = e.contacts[ci];
      Row r;
      so.side(r, 0, c.link_a, c.pos_a, c.normal, V3());
      so.side(r, 1, c.link_b, c.pos_b, -c.normal, V3());
      if (!so.finish(r)) continue;
      real rel = so.jv(r, so.vel);
      real pen = c.dist + (real)cfg.linear_slop;
      real poserr, velerr = -rel;
      if (pen > 0) { poserr = 0; velerr -= pen / dt; } else poserr = -pen * (real)cfg.contact_erp / dt;
      r.rhs = (poserr + velerr) * r.diag_inv; r.lo = 0; r.hi = real(1e30);
      if (wsc > 0) {                      // a persisting contact starts from its last normal impulse
        unsigned key = ckeys[ci];
        auto itp = std::lower_bound(e.ws_contacts.begin(), e.ws_contacts.end(), std::make_pair(key, real(-1e30)));
        if (itp != e.ws_contacts.end() && itp->first == key) r.lambda = wsc * itp->second;
      }
      crow[ci] = (int)rows.size();
      rows.push_back(r);
    }
    int first_friction_row = (int)rows.size();
    for (size_t ci = 0; ci < e.contacts.size(); ci++) {
      if (crow[ci] < 0) continue;
      Contact& c = e.contacts[ci];
      real mu = e.friction[c.link_a] * e.friction[c.link_b];
      V3 t1, t2; plane_space(c.normal, t1, t2);
      for (int d = 0; d < 2; d++) {
        V3 t = d == 0 ? t1 : t2;
        Row r;
        so.side(r, 0, c.link_a, c.pos_a, t, V3());
        so.side(r, 1, c.link_b, c.pos_b, -t, V3());
        if (!so.finish(r)) { r.n[0] = r.n[1] = 0; r.diag_inv = 0; r.lambda = 0; }
        r.rhs = (-so.jv(r, so.vel)) * r.diag_inv; r.lo = 0; r.hi = 0; r.friction_of = crow[ci]; r.mu = mu;
        rows.push_back(r);
      }
    }

    // 4. projected Gauss-Seidel on velocity deltas
    auto apply = [&](Row& r, real dl) {
      for (int sd = 0; sd < 2; sd++) for (int i = 0; i < r.n[sd]; i++) so.dv[r.off[sd] + i] += r.MiJ[sd][i] * dl;
    };
    for (int ri = 0; ri < first_friction_row; ri++) if (rows[ri].lambda != 0) apply(rows[ri], rows[ri].lambda);   // warm start
    int iters = 0;
    for (int it = 0; it < cfg.num_solver_iters; it++) {
      real resid = 0;
      iters = it + 1;
      for (int ri = 0; ri < first_friction_row; ri++) {
        Row& r = rows[ri];
        real dl = r.rhs - so.jv(r, so.dv) * r.diag_inv;
        real sum = r.lambda + dl;
        if (sum < r.lo) { dl = r.lo - r.lambda; sum = r.lo; } else if (sum > r.hi) { dl = r.hi - r.lambda; sum = r.hi; }
        r.lambda = sum; apply(r, dl);
        resid = std::max(resid, dl * dl);
      }
      for (int ri = first_friction_row; ri + 1 < (int)rows.size(); ri += 2) {
        Row& r1 = rows[ri]; Row& r2 = rows[ri + 1];
        real lim = r1.mu * rows[r1.friction_of].lambda;
        if (lim <= 0 && r1.lambda == 0 && r2.lambda == 0) continue;
        real d1 = r1.rhs - so.jv(r1, so.dv) * r1.diag_inv;
        real d2 = r2.rhs - so.jv(r2, so.dv) * r2.diag_inv;
        real s1 = r1.lambda + d1, s2 = r2.lambda + d2;
        if (cfg.cone_friction) {
          real mag2 = s1 * s1 + s2 * s2;
          if (mag2 > lim * lim) { real k = lim / std::sqrt(mag2); s1 *= k; s2 *= k; }
        } else {
          s1 = std::min(lim, std::max(-lim, s1)); s2 = std::min(lim, std::max(-lim, s2));
        }
        d1 = s1 - r1.lambda; d2 = s2 - r2.lambda;
        r1.lambda = s1; r2.lambda = s2;
        apply(r1, d1); apply(r2, d2);
        resid = std::max(resid, std::max(d1 * d1, d2 * d2));
      }
      if (cfg.residual_threshold > 0 && resid <= (real)cfg.residual_threshold) break;
    }
    e.last_iters = iters;

    // 5. write back velocities, integrate positions
    for (int b = 0; b < s.nb; b++) {
      if (so.body_off[b] < 0) continue;
      int o = so.body_off[b], l0 = s.body_link0[b];
      if (s.body_free[b]) {
        V3 v = e.base_lin[b], w = e.base_ang[b];
        for (int a = 0; a < 3; a++) {
          v[a] = std::min(vmax, std::max(-vmax, v[a] + so.dv[o + a]));
          w[a] = std::min(vmax, std::max(-vmax, w[a] + so.dv[o + 3 + a]));
        }
        e.base_lin[b] = v; e.base_ang[b] = w;
        V3 com = so.free_com[b] + v * dt;
        Quat qn = qnormalize(qmul(qexp(w * dt), e.base_quat[b]));
        e.base_quat[b] = qn;
        e.base_pos[b] = com - qrot(qn, s.link_com[l0]);
      } else {
        for (int k = l0 + 1; k < l0 + s.body_nlinks[b]; k++) if (s.link_live[k]) {
          real qd = std::min(vmax, std::max(-vmax, e.qd[k] + so.dv[o + s.link_dof[k]]));
          real qn = e.q[k] + dt * qd;
          if (e.hard_limit[k]) {               // Human.enforce_joint_limits (agent.py:240-250)
            if (qn < s.link_lower[k]) { qn = s.link_lower[k]; qd = 0; } else if (qn > s.link_upper[k]) { qn = s.link_upper[k]; qd = 0; }
          }
          e.qd[k] = qd; e.q[k] = qn;
        }
      }
    }
    for (int k = 0; k < s.nl; k++) if (motor_row[k] >= 0) e.motor_applied[k] = rows[motor_row[k]].lambda / dt;
    for (int k = 0; k < s.nl; k++) {
      e.ws_motor[k] = motor_row[k] >= 0 ? rows[motor_row[k]].lambda : 0;
      for (int sgn = 0; sgn < 2; sgn++) e.ws_limit[2 * k + sgn] = limit_row[2 * k + sgn] >= 0 ? rows[limit_row[2 * k + sgn]].lambda : 0;
    }
    for (int i = 0; i < 6 * s.ncon; i++) e.ws_fixed[i] = fixed_row[i] >= 0 ? rows[fixed_row[i]].lambda : 0;
    e.ws_contacts.clear();
    for (size_t ci = 0; ci < e.contacts.size(); ci++) if (crow[ci] >= 0) e.ws_contacts.push_back(std::make_pair(ckeys[ci], rows[crow[ci]].lambda));
    for (size_t ci = 0; ci < e.contacts.size(); ci++) {
      Contact& c = e.contacts[ci];
      if (crow[ci] < 0) continue;
      c.lambda_n = rows[crow[ci]].lambda;
    }
    int fr = first_friction_row;
    for (size_t ci = 0; ci < e.contacts.size(); ci++) {
      if (crow[ci] < 0) continue;
      e.contacts[ci].lambda_t1 = rows[fr].lambda; e.contacts[ci].lambda_t2 = rows[fr + 1].lambda; fr += 2;
    }
    (void)first_contact_row;
    // the cloth solves after the rigid world of this substep, against the START-of-substep poses (e.lpos, e.wverts: the
    // kinematics above ran before the integration)
    if (cloth) ocloth_substep(s, e, *cloth, *cloth_env, dt);
  }
  forward_kinematics(s, e);
}

void init_env(const Scene& s, Env& e) {
  e.base_pos.assign(s.nb, V3()); e.base_lin.assign(s.nb, V3()); e.base_ang.assign(s.nb, V3());
  e.base_quat.assign(s.nb, Quat());
  e.q.assign(s.nl, 0); e.qd.assign(s.nl, 0);
  e.motor_mode.assign(s.nl, AG_MOTOR_OFF); e.hard_limit.assign(s.nl, 0);
  e.motor_target.assign(s.nl, 0); e.motor_kp.assign(s.nl, 0); e.motor_kd.assign(s.nl, 0);
  e.motor_maxf.assign(s.nl, 0); e.motor_applied.assign(s.nl, 0); e.motor_fscale.assign(s.nl, 1);
  e.friction = s.link_friction;
  e.body_mode.assign(s.nb, 1);
  e.lpos.assign(s.nl, V3()); e.lquat.assign(s.nl, Quat());
  e.wverts.assign(s.nv, V3()); e.wplanes.assign(4 * s.np, 0);
  e.cmin.assign(s.nc, V3()); e.cmax.assign(s.nc, V3()); e.lmin.assign(s.nl, V3()); e.lmax.assign(s.nl, V3());
  e.overflow = 0; e.last_iters = 0;
}

inline bool mask_on(const int32_t* m, int i) { return !m || m[i]; }

}  // namespace

// ============================================================================ C API (double I/O)
extern "C" {

const char* oracle_last_error() { return g_err.c_str(); }

void oracle_default_config(AgConfig* c) {
  c->dt = 0.02; c->num_substeps = 1; c->num_solver_iters = 50; c->erp = 0.2; c->contact_erp = 0.08;
  c->warmstart_contact = 0.0; c->warmstart_joint = 0.0;
  c->linear_slop = 1e-5; c->residual_threshold = 1e-7; c->contact_threshold = 0.02;
  c->linear_damping = 0.04; c->angular_damping = 0.04; c->max_coord_velocity = 100; c->hull_margin = 0.001;
  c->cone_friction = 1; c->gyroscopic = 1; c->max_contacts = 128;
}

void* oracle_create(const AgSceneDesc* d, const AgConfig* cfg, int n_envs) {
  Sim* sim = new Sim();
  ingest(sim->sc, d);
  for (int b = 0; b < sim->sc.nb; b++) if (sim->sc.body_ndof[b] > MAXD) { g_err = "too many DoFs in one body"; delete sim; return nullptr; }
  sim->cfg = *cfg; sim->N = n_envs;
  sim->envs.resize(n_envs);
  for (auto& e : sim->envs) init_env(sim->sc, e);
  return sim;
}
void oracle_destroy(void* h) { delete (Sim*)h; }
int oracle_num_dofs(void* h, int body) { return ((Sim*)h)->sc.body_ndof[body]; }

int oracle_set_base_pose(void* h, int body, const double* pos, const double* quat, const int32_t* mask) {
  Sim* s = (Sim*)h;
  for (int i = 0; i < s->N; i++) if (mask_on(mask, i)) {
    if (pos) s->envs[i].base_pos[body] = V3((real)pos[3 * i], (real)pos[3 * i + 1], (real)pos[3 * i + 2]);
    if (quat) s->envs[i].base_quat[body] = qnormalize(Quat((real)quat[4 * i], (real)quat[4 * i + 1], (real)quat[4 * i + 2], (real)quat[4 * i + 3]));
  }
  return 0;
}
int oracle_set_base_velocity(void* h, int body, const double* lin, const double* ang, const int32_t* mask) {
  Sim* s = (Sim*)h;
  for (int i = 0; i < s->N; i++) if (mask_on(mask, i)) {
    if (lin) s->envs[i].base_lin[body] = V3((real)lin[3 * i], (real)lin[3 * i + 1], (real)lin[3 * i + 2]);
    if (ang) s->envs[i].base_ang[body] = V3((real)ang[3 * i], (real)ang[3 * i + 1], (real)ang[3 * i + 2]);
  }
  return 0;
}
int oracle_set_joint_state(void* h, int n, const int32_t* links, const double* q, const double* qd, const int32_t* mask) {
  Sim* s = (Sim*)h;
  for (int i = 0; i < s->N; i++) if (mask_on(mask, i)) for (int j = 0; j < n; j++) {
    if (q) s->envs[i].q[links[j]] = (real)q[i * n + j];
    if (qd) s->envs[i].qd[links[j]] = (real)qd[i * n + j];
  }
  return 0;
}
int oracle_set_link_friction(void* h, int link, const double* mu, const int32_t* mask) {
  Sim* s = (Sim*)h;
  for (int i = 0; i < s->N; i++) if (mask_on(mask, i)) s->envs[i].friction[link] = (real)mu[i];
  return 0;
}
int oracle_set_body_mode(void* h, int body, const int32_t* mode) {
  Sim* s = (Sim*)h;
  for (int i = 0; i < s->N; i++) s->envs[i].body_mode[body] = mode[i];
  return 0;
}
int oracle_set_hard_limits(void* h, int n, const int32_t* links, int on) {
  Sim* s = (Sim*)h;
  for (auto& e : s->envs) for (int j = 0; j < n; j++) e.hard_limit[links[j]] = on ? 1 : 0;
  return 0;
}
int oracle_set_motor(void* h, int n, const int32_t* links, int mode, const double* target, const double* kp,
                     const double* kd, const double* maxf) {
  Sim* s = (Sim*)h;
  for (int i = 0; i < s->N; i++) for (int j = 0; j < n; j++) {
    Env& e = s->envs[i]; int k = links[j];
    e.motor_mode[k] = mode;
    if (target) e.motor_target[k] = (real)target[i * n + j];
    if (kp) e.motor_kp[k] = (real)kp[j];
    if (kd) e.motor_kd[k] = (real)kd[j];
    if (maxf) e.motor_maxf[k] = (real)maxf[j];
  }
  return 0;
}
int oracle_set_motor_force_scale(void* h, int n, const int32_t* links, const double* scale) {
  Sim* s = (Sim*)h;
  for (int i = 0; i < s->N; i++) for (int j = 0; j < n; j++) s->envs[i].motor_fscale[links[j]] = (real)scale[i * n + j];
  return 0;
}
int oracle_set_motor_targets(void* h, int n, const int32_t* links, const double* target) {
  Sim* s = (Sim*)h;
  for (int i = 0; i < s->N; i++) for (int j = 0; j < n; j++) s->envs[i].motor_target[links[j]] = (real)target[i * n + j];
  return 0;
}
int oracle_set_body_gravity(void* h, int body, const double* g) { ((Sim*)h)->sc.body_gravity[body] = V3((real)g[0], (real)g[1], (real)g[2]); return 0; }
int oracle_forward_kinematics(void* h) {
  Sim* s = (Sim*)h;
  for (auto& e : s->envs) forward_kinematics(s->sc, e);
  return 0;
}
int oracle_step(void* h, int n_steps, int n_threads) {
  Sim* s = (Sim*)h;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads > 0 ? n_threads : 1)
#endif
  for (int i = 0; i < s->N; i++)
    for (int k = 0; k < n_steps; k++) step_env(s->sc, s->cfg, s->envs[i], s->has_cloth ? &s->cloth : nullptr, s->has_cloth ? &s->cloth_envs[i] : nullptr);
  (void)n_threads;
  return 0;
}
int oracle_get_joint_states(void* h, int n, const int32_t* links, double* q, double* qd, double* tau) {
  Sim* s = (Sim*)h;
  for (int i = 0; i < s->N; i++) for (int j = 0; j < n; j++) {
    const Env& e = s->envs[i]; int k = links[j];
    if (q) q[i * n + j] = e.q[k];
    if (qd) qd[i * n + j] = e.qd[k];
    if (tau) tau[i * n + j] = e.motor_applied[k];
  }
  return 0;
}
int oracle_get_link_states(void* h, int n, const int32_t* links, double* pos, double* quat, double* com_pos,
                           double* com_quat, double* lin_vel, double* ang_vel) {
  Sim* s = (Sim*)h; const Scene& sc = s->sc;
  std::vector<SV> v(sc.nl);
  for (int i = 0; i < s->N; i++) {
    const Env& e = s->envs[i];
    for (int b = 0; b < sc.nb; b++) {
      int l0 = sc.body_link0[b];
      if (sc.body_free[b]) {
        V3 com = world_com(sc, e, l0);
        v[l0] = SV(e.base_ang[b], e.base_lin[b] - cross(e.base_ang[b], com));
      } else {
        std::vector<SV> tmp(sc.nl);
        link_velocities(sc, e, b, tmp);
        for (int k = l0; k < l0 + sc.body_nlinks[b]; k++) v[k] = tmp[k];
      }
    }
    for (int j = 0; j < n; j++) {
      int k = links[j]; int o = i * n + j;
      V3 com = world_com(sc, e, k);
      Quat cq = qmul(e.lquat[k], sc.link_iquat[k]);
      if (pos) { pos[3 * o] = e.lpos[k].x; pos[3 * o + 1] = e.lpos[k].y; pos[3 * o + 2] = e.lpos[k].z; }
      if (quat) { quat[4 * o] = e.lquat[k].x; quat[4 * o + 1] = e.lquat[k].y; quat[4 * o + 2] = e.lquat[k].z; quat[4 * o + 3] = e.lquat[k].w; }
      if (com_pos) { com_pos[3 * o] = com.x; com_pos[3 * o + 1] = com.y; com_pos[3 * o + 2] = com.z; }
      if (com_quat) { com_quat[4 * o] = cq.x; com_quat[4 * o + 1] = cq.y; com_quat[4 * o + 2] = cq.z; com_quat[4 * o + 3] = cq.w; }
      V3 lv = v[k].l + cross(v[k].a, com);
      if (lin_vel) { lin_vel[3 * o] = lv.x; lin_vel[3 * o + 1] = lv.y; lin_vel[3 * o + 2] = lv.z; }
      if (ang_vel) { ang_vel[3 * o] = v[k].a.x; ang_vel[3 * o + 1] = v[k].a.y; ang_vel[3 * o + 2] = v[k].a.z; }
    }
  }
  return 0;
}

static void fill_contact(const Sim* s, const Contact& c, bool flip, double dt, AgContact* o) {
  (void)s;
  o->link_a = flip ? c.link_b : c.link_a; o->link_b = flip ? c.link_a : c.link_b;
  V3 pa = flip ? c.pos_b : c.pos_a, pb = flip ? c.pos_a : c.pos_b, n = flip ? -c.normal : c.normal;
  for (int a = 0; a < 3; a++) { o->pos_a[a] = (float)pa[a]; o->pos_b[a] = (float)pb[a]; o->normal[a] = (float)n[a]; }
  o->distance = (float)c.dist; o->normal_force = (float)(c.lambda_n / dt);
}
static bool link_match(const Scene& sc, int link, int body, int lidx) {
  if (sc.link_body[link] != body) return false;
  if (lidx == -2) return true;
  return link == sc.body_link0[body] + 1 + lidx;
}
int oracle_get_contacts(void* h, int body_a, int body_b, int link_a, int link_b, int max_pts, AgContact* out, int32_t* count) {
  Sim* s = (Sim*)h; const Scene& sc = s->sc;
  double dt = s->cfg.dt / std::max(1, s->cfg.num_substeps);
  for (int i = 0; i < s->N; i++) {
    int n = 0;
    for (const Contact& c : s->envs[i].contacts) {
      bool fwd = link_match(sc, c.link_a, body_a, link_a) && (body_b == -2 || link_match(sc, c.link_b, body_b, link_b));
      bool rev = link_match(sc, c.link_b, body_a, link_a) && (body_b == -2 || link_match(sc, c.link_a, body_b, link_b));
      if (!fwd && !rev) continue;
      if (n < max_pts && out) fill_contact(s, c, !fwd, dt, &out[(size_t)i * max_pts + n]);
      n++;
    }
    count[i] = n;
  }
  return 0;
}
int oracle_contact_force_sum(void* h, int body_a, int body_b, int link_a, int link_b, double* out) {
  Sim* s = (Sim*)h; const Scene& sc = s->sc;
  double dt = s->cfg.dt / std::max(1, s->cfg.num_substeps);
  for (int i = 0; i < s->N; i++) {
    double f = 0;
    for (const Contact& c : s->envs[i].contacts) {
      bool fwd = link_match(sc, c.link_a, body_a, link_a) && (body_b == -2 || link_match(sc, c.link_b, body_b, link_b));
      bool rev = link_match(sc, c.link_b, body_a, link_a) && (body_b == -2 || link_match(sc, c.link_a, body_b, link_b));
      if (fwd || rev) f += c.lambda_n / dt;
    }
    out[i] = f;
  }
  return 0;
}
int oracle_closest_points(void* h, int body_a, int body_b, double distance, int max_pts, AgContact* out, int32_t* count) {
  Sim* s = (Sim*)h; const Scene& sc = s->sc;
  for (int i = 0; i < s->N; i++) {
    Env& e = s->envs[i];
    forward_kinematics(sc, e);
    update_colliders(sc, e);
    int n = 0;
    if (e.body_mode[body_a] == 0 || e.body_mode[body_b] == 0) { count[i] = 0; continue; }
    int a0 = sc.body_link0[body_a], b0 = sc.body_link0[body_b];
    for (int la = a0; la < a0 + sc.body_nlinks[body_a]; la++)
      for (int lb = b0; lb < b0 + sc.body_nlinks[body_b]; lb++) {
        if (!sc.link_ncol[la] || !sc.link_ncol[lb]) continue;
        if (!aabb_overlap(e.lmin[la], e.lmax[la], e.lmin[lb], e.lmax[lb], (real)distance)) continue;
        for (int ca = sc.link_col0[la]; ca < sc.link_col0[la] + sc.link_ncol[la]; ca++)
          for (int cb = sc.link_col0[lb]; cb < sc.link_col0[lb] + sc.link_ncol[lb]; cb++) {
            if (!aabb_overlap(e.cmin[ca], e.cmax[ca], e.cmin[cb], e.cmax[cb], (real)distance)) continue;
            Contact c[4];
            if (!collide_pair(sc, e, ca, cb, (real)distance, false, c)) continue;
            if (n < max_pts && out) fill_contact(s, c[0], false, 1.0, &out[(size_t)i * max_pts + n]);
            n++;
          }
      }
    count[i] = n;
  }
  return 0;
}
int oracle_num_contacts(void* h, int32_t* count, int32_t* iters) {
  Sim* s = (Sim*)h;
  for (int i = 0; i < s->N; i++) { if (count) count[i] = (int)s->envs[i].contacts.size(); if (iters) iters[i] = s->envs[i].last_iters; }
  return 0;
}

// state blob: per body pos3 quat4 lin3 ang3, per link q qd
size_t oracle_state_size(void* h) { Sim* s = (Sim*)h; return (size_t)s->sc.nb * 13 + (size_t)s->sc.nl * 2; }
int oracle_state_get(void* h, double* out) {
  Sim* s = (Sim*)h; size_t sz = oracle_state_size(h);
  for (int i = 0; i < s->N; i++) {
    const Env& e = s->envs[i]; double* o = out + sz * i;
    for (int b = 0; b < s->sc.nb; b++) {
      for (int a = 0; a < 3; a++) { o[a] = e.base_pos[b][a]; o[7 + a] = e.base_lin[b][a]; o[10 + a] = e.base_ang[b][a]; }
      o[3] = e.base_quat[b].x; o[4] = e.base_quat[b].y; o[5] = e.base_quat[b].z; o[6] = e.base_quat[b].w;
      o += 13;
    }
    for (int k = 0; k < s->sc.nl; k++) { o[0] = e.q[k]; o[1] = e.qd[k]; o += 2; }
  }
  return 0;
}
int oracle_state_set(void* h, const double* in) {
  Sim* s = (Sim*)h; size_t sz = oracle_state_size(h);
  for (int i = 0; i < s->N; i++) {
    Env& e = s->envs[i]; const double* o = in + sz * i;
    for (int b = 0; b < s->sc.nb; b++) {
      for (int a = 0; a < 3; a++) { e.base_pos[b][a] = (real)o[a]; e.base_lin[b][a] = (real)o[7 + a]; e.base_ang[b][a] = (real)o[10 + a]; }
      e.base_quat[b] = Quat((real)o[3], (real)o[4], (real)o[5], (real)o[6]);
      o += 13;
    }
    for (int k = 0; k < s->sc.nl; k++) { e.q[k] = (real)o[0]; e.qd[k] = (real)o[1]; o += 2; }
    forward_kinematics(s->sc, e);
  }
  return 0;
}

// diagnostics used by unit tests: generalized inverse mass matrix of a fixed-base body (CRBA route)

// ---- cloth (mirrors ag_cloth_* of include/agphys.h, double I/O)
int oracle_cloth_init(void* h, const AgClothDesc* d) {
  Sim* s = (Sim*)h;
  OCloth& C = s->cloth;
  C.nn = d->n_nodes; C.piters = d->piterations; C.maxcc = d->max_contacts > 0 ? d->max_contacts : 1024;
  C.links.assign(d->links, d->links + 2 * d->n_links);
  C.rest2.resize(d->n_links); for (int l = 0; l < d->n_links; l++) C.rest2[l] = (real)d->link_rest2[l];
  C.nf_off.assign(d->nf_off, d->nf_off + d->n_nodes + 1); C.nf_pair.assign(d->nf_pair, d->nf_pair + 2 * d->n_nf);
  C.area.resize(d->n_nodes); for (int i = 0; i < d->n_nodes; i++) C.area[i] = (real)d->node_area[i];
  C.im = (real)d->inv_mass; C.kLST = (real)d->kLST; C.kDP = (real)d->kDP; C.kDG = (real)d->kDG; C.kLF = (real)d->kLF; C.kDF = (real)d->kDF;
  C.kCHR = (real)d->kCHR; C.kKHR = (real)d->kKHR; C.kAHR = (real)d->kAHR; C.margin = (real)d->margin; C.density = (real)d->air_density;
  C.gravity = V3((real)d->gravity[0], (real)d->gravity[1], (real)d->gravity[2]);
  C.anchor_node.assign(d->anchor_node, d->anchor_node + d->n_anchors);
  C.anchor_local.clear(); for (int a = 0; a < d->n_anchors; a++) C.anchor_local.push_back(rd3(d->anchor_local, a));
  C.col_links.assign(d->col_links, d->col_links + d->n_col_links); C.col_static.assign(d->col_link_static, d->col_link_static + d->n_col_links);
  C.bs_c.clear(); C.bs_r.clear();
  for (int L = 0; L < d->n_col_links; L++) { C.bs_c.push_back(V3((real)d->col_link_bsphere[4 * L], (real)d->col_link_bsphere[4 * L + 1], (real)d->col_link_bsphere[4 * L + 2])); C.bs_r.push_back((real)d->col_link_bsphere[4 * L + 3]); }
  s->cloth_envs.assign(s->N, OClothEnv());
  for (auto& ce : s->cloth_envs) { ce.x.assign(C.nn, V3()); ce.v.assign(C.nn, V3()); ce.q.assign(C.nn, V3()); }
  s->has_cloth = true;
  return 0;
}
int oracle_cloth_set_state(void* h, const double* x, const double* v, const int32_t* mask) {
  Sim* s = (Sim*)h; int nn = s->cloth.nn;
  for (int e = 0; e < s->N; e++) if (mask_on(mask, e)) for (int i = 0; i < nn; i++) {
    if (x) s->cloth_envs[e].x[i] = rd3(x, e * nn + i);
    if (v) s->cloth_envs[e].v[i] = rd3(v, e * nn + i);
  }
  return 0;
}
int oracle_cloth_get_state(void* h, double* x, double* v) {
  Sim* s = (Sim*)h; int nn = s->cloth.nn;
  for (int e = 0; e < s->N; e++) for (int i = 0; i < nn; i++) for (int c = 0; c < 3; c++) {
    if (x) x[((size_t)e * nn + i) * 3 + c] = s->cloth_envs[e].x[i][c];
    if (v) v[((size_t)e * nn + i) * 3 + c] = s->cloth_envs[e].v[i][c];
  }
  return 0;
}
int oracle_cloth_set_anchor(void* h, const double* pos, const int32_t* mask) {
  Sim* s = (Sim*)h;
  for (int e = 0; e < s->N; e++) if (mask_on(mask, e)) s->cloth_envs[e].anchor_pos = rd3(pos, e);
  return 0;
}
int oracle_cloth_anchor_follow(void* h, int link) {
  Sim* s = (Sim*)h;
  for (int e = 0; e < s->N; e++) s->cloth_envs[e].anchor_pos = s->envs[e].lpos[link];
  return 0;
}
int oracle_cloth_set_gravity(void* h, const double* g) { ((Sim*)h)->cloth.gravity = V3((real)g[0], (real)g[1], (real)g[2]); return 0; }
int oracle_cloth_get_contacts(void* h, int max_pts, int32_t* count, int32_t* node, double* pos, double* force, int32_t* link) {
  Sim* s = (Sim*)h;
  double dt = s->cfg.dt / std::max(1, s->cfg.num_substeps);
  double fs = 1.0 / ((double)s->cloth.im * dt * dt);
  for (int e = 0; e < s->N; e++) {
    const OClothEnv& ce = s->cloth_envs[e];
    if (count) count[e] = (int)ce.contacts.size();
    for (int k = 0; k < std::min((int)ce.contacts.size(), max_pts); k++) {
      const OClothContact& c = ce.contacts[k];
      size_t o = (size_t)e * max_pts + k;
      if (node) node[o] = c.node;
      if (link) link[o] = c.link;
      for (int a = 0; a < 3; a++) { if (pos) pos[3 * o + a] = ce.x[c.node][a]; if (force) force[3 * o + a] = -(double)c.acc[a] * fs; }
    }
  }
  return 0;
}

int oracle_mass_matrix_inv(void* h, int env, int body, double* out) {
  Sim* s = (Sim*)h; Env& e = s->envs[env];
  forward_kinematics(s->sc, e);
  BodyDyn bd; crba_factor(s->sc, e, body, bd);
  int n = bd.ndof;
  for (int j = 0; j < n; j++) {
    real rhs[MAXD], x[MAXD];
    for (int i = 0; i < n; i++) rhs[i] = (i == j);
    chol_solve(bd, rhs, x);
    for (int i = 0; i < n; i++) out[i * n + j] = x[i];
  }
  return n;
}
// standalone GJK entry for unit tests (vertex sets in world coordinates)
int oracle_gjk(const double* A, int nA, const double* B, int nB, double* pa, double* pb, double* dist) {
  std::vector<V3> a(nA), b(nB);
  for (int i = 0; i < nA; i++) a[i] = V3((real)A[3 * i], (real)A[3 * i + 1], (real)A[3 * i + 2]);
  for (int i = 0; i < nB; i++) b[i] = V3((real)B[3 * i], (real)B[3 * i + 1], (real)B[3 * i + 2]);
  ClosestResult r = gjk_closest(a.data(), nA, b.data(), nB);
  for (int k = 0; k < 3; k++) { pa[k] = r.pa[k]; pb[k] = r.pb[k]; }
  *dist = r.dist;
  return r.overlap ? 1 : 0;
}

}  // extern "C"

// diagnostics: narrowphase work per enabled link pair for one env (number of collider pairs that
// survive the AABB culls, and the vertex-pair product they represent)
extern "C" int oracle_debug_pair_work(void* h, int env, int32_t* n_narrow, double* vert_work) {
  Sim* s = (Sim*)h; const Scene& sc = s->sc; Env& e = s->envs[env];
  forward_kinematics(sc, e); update_colliders(sc, e);
  real fac = (real)s->cfg.contact_threshold;
  for (int p = 0; p < sc.npair; p++) {
    n_narrow[p] = 0; vert_work[p] = 0;
    int la = sc.pair_link[2 * p], lb = sc.pair_link[2 * p + 1];
    if (e.body_mode[sc.link_body[la]] == 0 || e.body_mode[sc.link_body[lb]] == 0) continue;
    if (!aabb_overlap(e.lmin[la], e.lmax[la], e.lmin[lb], e.lmax[lb], fac * std::min(sc.link_thresh[la], sc.link_thresh[lb]))) continue;
    n_narrow[p] = -1;   // link-level overlap, maybe no collider pair
    int cnt = 0;
    for (int ca = sc.link_col0[la]; ca < sc.link_col0[la] + sc.link_ncol[la]; ca++) {
      if (!aabb_overlap(e.cmin[ca], e.cmax[ca], e.lmin[lb], e.lmax[lb], fac * sc.col_thresh[ca])) continue;
      for (int cb = sc.link_col0[lb]; cb < sc.link_col0[lb] + sc.link_ncol[lb]; cb++) {
        real thr = fac * std::min(sc.col_thresh[ca], sc.col_thresh[cb]);
        if (!aabb_overlap(e.cmin[ca], e.cmax[ca], e.cmin[cb], e.cmax[cb], thr)) continue;
        cnt++; vert_work[p] += (double)(sc.col_nv[ca] + sc.col_nv[cb]);
      }
    }
    if (cnt) n_narrow[p] = cnt;
  }
  return 0;
}
