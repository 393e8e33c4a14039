// ag_solver.cuh — K6 (constraint rows -> packed per-env row stream) and K7 (PGS over the stream).
//
// What it restates: Bullet's btMultiBodyConstraintSolver for one p.stepSimulation (envs/env.py:226):
// joint-limit rows (only while violated), joint-motor rows (agent.py:33), the spoon<->gripper fixed
// constraint (tool.py:46-47, 6 rows, maxForce 500) and frictional contacts (1 normal + 2 friction
// rows, implicit cone), solved by projected Gauss-Seidel in that order, 50 iterations, early exit on
// the least-squares residual.
//
// B200 design.  The Gauss-Seidel chain of one env is strictly sequential, so K7's speed is the latency
// and the instruction count of one row update.  K6 therefore writes every row as a self-contained
// record -- both sides' Jacobian entries J and M^-1 J^T, rhs, 1/(J M^-1 J^T), bounds -- into ONE
// contiguous, env-major stream of 128-byte slots in HBM/L2.  K7 keeps only the mutable state (velocity
// deltas, impulses: ~2 KB/env) in shared memory, so all 4096 envs are resident at once, and pulls
// the read-only stream through a two-deep ring of 1 KB buffers with TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx): the next 8 slots stream in from L2 while the current 8 are
// being solved.  A row update is then vector LDS of the record + the two bodies' velocities, 12-24 FMA,
// a clamp, and vector STS.  Records never straddle a 1 KB chunk (the packer pads).
#pragma once
#include "ag_device.cuh"

#define RS_SLOT 32            // floats per slot (128 B)
#define RS_CHUNK 8            // slots per TMA chunk (1 KB)
// record codes.  F = one free body against something static, FF = two free bodies, GEN = any articulated side
enum { RK_ROW_F = 0, RK_FRIC_F = 1, RK_ROW_FF = 2, RK_FRIC_FF = 3, RK_ROW_GEN = 4, RK_FRIC_GEN = 5, RK_PAD = 7 };
// Record layouts (floats; every record starts with [0] code | nslots << 4, [1] offA | nA << 16, [3] impulse index):
//   ROW_F    [4..7] rhs dinv lo hi, [8..13] J_A, [16..21] M_A
//   ROW_FF   [2] offB, [4..7], [8..13] J_A [14..19] J_B, [20..25] M_A [26..31] M_B
//   ROW_GEN  [2] offB | nB << 16, [4..7], J at 8 (side A padded to 4, then side B padded to 4), M at 8 + P;
//            side A is always an articulation (limit / motor rows: J = +-e_d), side B nothing, a free body or a second articulation
//   FRIC_F   [2] mu, [4..7] rhs1 dinv1 rhs2 dinv2, [8..13] J1 [14..19] J2 [20..25] M1 [26..31] M2
//   FRIC_FF  [2] offB, [4..7], [8..19] J1 (A,B) [20..31] J2 | second slot: [32..43] M1 [44..55] M2 [56] mu
//   FRIC_GEN [2] offB | nB << 16, [4..7], [8] mu, J1 at 12, J2 at 12 + P, M1 at 12 + 2P, M2 at 12 + 3P
// A free body's 6 entries are (lin xyz, ang xyz) and match its 8-float block of the velocity vector.

struct alignas(16) v4 { float x, y, z, w; };
struct alignas(8) v2 { float x, y; };
AG_HD v4 ldv4(const float* p) { return *(const v4*)p; }
AG_HD v2 ldv2(const float* p) { return *(const v2*)p; }
AG_HD void stv4(float* p, v4 a) { *(v4*)p = a; }
AG_HD void stv2(float* p, v2 a) { *(v2*)p = a; }
AG_HD float i2f_bits(int v) { float f; memcpy(&f, &v, 4); return f; }
AG_HD int f2i_bits(float f) { int v; memcpy(&v, &f, 4); return v; }
AG_HD float dot4(v4 a, v4 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }

// velocity-delta vector of one env: [articulation a: its dofs, padded to 4, at art_voff[a]]...[free body f: lin xyz, ang xyz, 2 pad]
AG_HD int rs_nv(const SimDev& S) { return S.NDp + 8 * S.nf; }
// impulses: [3 ND dof rows][ngr fixed-constraint rows][3 per contact]
AG_HD int rs_nlam(const SimDev& S) { return (3 * S.ND + S.ngr + 3 * S.maxc + 3) & ~3; }
AG_HD int rs_lane_floats(const SimDev& S) {
  int t = rs_nv(S) + rs_nlam(S) + S.rs_nbuf * RS_CHUNK * RS_SLOT;
  while (t % 32 != 8) t += 4;                    // lanes of a CTA start 8 banks apart
  return t;
}
AG_HD int rs_pad4(int n) { return (n + 3) & ~3; }
AG_HD int rs_slots(int floats) { return (floats + RS_SLOT - 1) / RS_SLOT; }

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

// where a side's velocity entries live in the env's velocity-delta vector, and how many there are
AG_HD void side_dims(const SimDev& S, int ref, int& off, int& n) {
  int kind = ref & 3, idx = ref >> 2;
  if (kind == 1) { off = S.NDp + 8 * idx; n = 6; }
  else if (kind == 2) { int a = AG_LDG(S.dl_art + idx); off = AG_LDG(S.art_voff + a); n = AG_LDG(S.art_nd + a); }
  else { off = 0; n = 0; }
}
// side order of a record: a lone dynamic side is A, an articulation paired with a free body is A (J carries the sign)
AG_HD bool rs_swap_sides(int refA, int refB) {
  int ka = refA & 3, kb = refB & 3;
  return (ka == 0 && kb != 0) || (ka == 1 && kb == 2);
}
// record code and slot count of a single row / a friction pair between sides of nA, nB entries (kinds from the refs)
AG_HD int rs_row_code(int refA, int refB, int nA, int nB, int& ns) {
  int ka = refA & 3, kb = refB & 3;
  if (ka == 1 && kb == 0) { ns = 1; return RK_ROW_F; }
  if (ka == 1 && kb == 1) { ns = 1; return RK_ROW_FF; }
  ns = rs_slots(8 + 2 * (rs_pad4(nA) + rs_pad4(nB)));
  return RK_ROW_GEN;
}
AG_HD int rs_fric_code(int refA, int refB, int nA, int nB, int& ns) {
  int ka = refA & 3, kb = refB & 3;
  if (ka == 1 && kb == 0) { ns = 1; return RK_FRIC_F; }
  if (ka == 1 && kb == 1) { ns = 2; return RK_FRIC_FF; }
  ns = rs_slots(12 + 4 * (rs_pad4(nA) + rs_pad4(nB)));
  return RK_FRIC_GEN;
}

// One side of a row: unit force `lin` at world point p plus torque `ang`.  Writes the side's J entries to
// Jd[0..n) and M^-1 J^T to Md[0..n) (zero-filled up to `pad`), returns J M^-1 J^T and accumulates J.v into rel.
AG_HDN inline float emit_side(const SimDev& S, int e, int ref, f3 p, f3 lin, f3 ang, float* Jd, float* Md, int pad, float& rel) {
  const int N = S.N;
  int kind = ref & 3, idx = ref >> 2;
  float diag = 0.f;
  int n = 0;
  if (kind == 1) {
    int b = AG_LDG(S.free_body + idx);
    float invm = AG_LDG(S.free_invm + idx);
    f3 r = p - ld3(S.fcom, idx, N, e);
    f3 t = cross(r, lin) + ang;
    f3 it = mul(ld_Iinv(S, idx, e), t);
    f3 v = ld3(S.base_lin, b, N, e), w = ld3(S.base_ang, b, N, e);
    rel += dot(lin, v) + dot(t, w);
    Jd[0] = lin.x; Jd[1] = lin.y; Jd[2] = lin.z; Jd[3] = t.x; Jd[4] = t.y; Jd[5] = t.z;
    Md[0] = lin.x * invm; Md[1] = lin.y * invm; Md[2] = lin.z * invm; Md[3] = it.x; Md[4] = it.y; Md[5] = it.z;
    diag = invm * dot(lin, lin) + dot(t, it);
    n = 6;
  } else if (kind == 2) {
    float J[AG_MAXND];
    int a = AG_LDG(S.dl_art + idx), d0 = AG_LDG(S.art_dl0 + a), nd = AG_LDG(S.art_nd + a);
    for (int i = 0; i < nd; i++) J[i] = 0.f;
    int j = idx;
    while (j >= 0) {
      f3 axw = ld3(S.jax, j, N, e), o = ld3(S.jor, j, N, e);
      J[j - d0] = (AG_LDG(S.dl_type + j) == 1) ? (dot(lin, cross(axw, p - o)) + dot(ang, axw)) : dot(lin, axw);
      j = AG_LDG(S.dl_parent + j);
    }
    for (int i = 0; i < nd; i++) {
      float m = 0.f;
      for (int k = 0; k < nd; k++) m += S.Minv[((size_t)(d0 + i) * S.ND + (d0 + k)) * N + e] * J[k];
      Jd[i] = J[i]; Md[i] = m;
      diag += J[i] * m;
      rel += J[i] * ld1(S.jqd, AG_LDG(S.dl_link + d0 + i), N, e);
    }
    n = nd;
  }
  for (int i = n; i < pad; i++) { Jd[i] = 0.f; Md[i] = 0.f; }
  return diag;
}

// reserve `ns` slots of the env's stream; a record never straddles a chunk (pad record up to the boundary)
AG_HD int rs_alloc(int& pos, int ns, float* rs, int cap) {
  int room = RS_CHUNK - (pos % RS_CHUNK);
  if (ns > room) {
    if (pos + room > cap) return -1;
    rs[(size_t)pos * RS_SLOT] = i2f_bits(RK_PAD | (room << 4));
    pos += room;
  }
  if (pos + ns > cap) return -1;
  int o = pos; pos += ns;
  return o;
}
AG_HD void rs_header(float* d, int code, int ns, int offA, int nA, int offB, int nB, int lam) {
  d[0] = i2f_bits(code | (ns << 4)); d[1] = i2f_bits(offA | (nA << 16)); d[2] = i2f_bits(offB | (nB << 16)); d[3] = i2f_bits(lam);
}

// Is dof row r = kind * ND + d (kind 0 lower limit, 1 upper limit, 2 motor) live this substep?  If so: its constants.
struct DofRow { float rhs, dinv, lo, hi, sgn; };
AG_HD bool dof_row(const SimDev& S, int e, int kind, int d, DofRow& R) {
  const int N = S.N;
  const float dt = S.dt;
  int k = AG_LDG(S.dl_link + d);
  float Mdd = S.Minv[((size_t)d * S.ND + d) * N + e];
  if (!(Mdd > 0.f)) return false;
  R.dinv = 1.0f / Mdd; R.lo = 0.f; R.hi = 1e30f; R.sgn = 1.f;
  float q = ld1(S.jq, k, N, e), qd = ld1(S.jqd, k, N, e);
  if (kind < 2) {             // limits: a row only while violated
    if (!AG_LDG(S.link_haslimit + k)) return false;
    if (kind == 0) { float pen = q - AG_LDG(S.link_lower + k); if (pen > 0.f) return false; R.rhs = (-pen * S.erp / dt - qd) * R.dinv; }
    else { float pen = AG_LDG(S.link_upper + k) - q; if (pen > 0.f) return false; R.rhs = (-pen * S.erp / dt + qd) * R.dinv; R.sgn = -1.f; }
  } else {
    int mode = S.motor_mode[k];
    float maxi = S.motor_maxf[k] * dt;
    if (mode == 0 || !(maxi > 0.f)) return false;
    float vt = (mode == 1) ? (S.motor_kp[k] * (ld1(S.motor_target, k, N, e) - q) / dt + qd - S.motor_kd[k] * qd)
                           : ld1(S.motor_target, k, N, e);
    R.rhs = (vt - qd) * R.dinv; R.lo = -maxi; R.hi = maxi;
  }
  return true;
}
// sides of fixed constraint c in record order (articulation / lone dynamic side first)
AG_HD bool con_sides(const SimDev& S, int e, int c, int& refA, int& refB, bool& swapped) {
  const int N = S.N;
  int ka = AG_LDG(S.con_link + 2 * c), kb = AG_LDG(S.con_link + 2 * c + 1);
  int ba = AG_LDG(S.link_body + ka), bb = AG_LDG(S.link_body + kb);
  if (S.body_mode[(size_t)ba * N + e] == 0 || S.body_mode[(size_t)bb * N + e] == 0) return false;
  refA = link_ref(S, e, ka); refB = link_ref(S, e, kb);
  swapped = rs_swap_sides(refA, refB);
  if (swapped) { int t = refA; refA = refB; refB = t; }
  return ((refA | refB) & 3) != 0;
}

// Sides of contact `key` in record order; called by K4 for every sorted contact (one thread each) so that K6a's
// sequential pass only reads two ints per contact.
AG_HD void contact_refs(const SimDev& S, int e, unsigned key, int& refA, int& refB) {
  unsigned pairk = key >> 2;
  int ca = (int)(pairk / (unsigned)S.nc), cb = (int)(pairk % (unsigned)S.nc);
  refA = link_ref(S, e, AG_LDG(S.col_link + ca)); refB = link_ref(S, e, AG_LDG(S.col_link + cb));
  if (rs_swap_sides(refA, refB)) { int t = refA; refA = refB; refB = t | (1 << 30); }   // bit 30 of refB: sides were swapped
}

// K6a: one lane per env: the slot layout of this substep's row stream in solver order -- joint-limit rows,
// motor rows, fixed-constraint rows, contact normal rows (contact order), friction pairs (contact order).
// Cheap and sequential; the records themselves are written by K6b with one thread per row.
// row_off [3 ND + ngr][N]: slot of each dof / fixed-constraint row (-1: not live); s_ref[..][2,3]: contact rows.
AG_HDN inline void rows_body(int e, const SimDev& S, const KP&) {
  const int N = S.N;
  float* rs = S.rs_data + (size_t)e * S.rs_cap * RS_SLOT;
  const int cap = S.rs_cap;
  int pos = 0;
  bool over = false;
  for (int kind = 0; kind < 3; kind++) {
    for (int d = 0; d < S.ND; d++) {
      DofRow R;
      int o = -1;
      if (dof_row(S, e, kind, d, R)) {
        int nd4 = rs_pad4(AG_LDG(S.art_nd + AG_LDG(S.dl_art + d)));
        o = rs_alloc(pos, rs_slots(8 + 2 * nd4), rs, cap);
        if (o < 0) over = true;
      }
      S.row_off[(size_t)(kind * S.ND + d) * N + e] = o;
    }
  }
  for (int c = 0; c < S.ncon; c++) {
    int refA, refB; bool sw;
    bool on = con_sides(S, e, c, refA, refB, sw);
    int offA, nA, offB, nB, ns = 0;
    if (on) { side_dims(S, refA, offA, nA); side_dims(S, refB, offB, nB); rs_row_code(refA, refB, nA, nB, ns); }
    for (int i = 0; i < 6; i++) {
      int o = -1;
      if (on) { o = rs_alloc(pos, ns, rs, cap); if (o < 0) over = true; }
      S.row_off[(size_t)(3 * S.ND + 6 * c + i) * N + e] = o;
    }
  }
  int cnt = S.c_count[e]; if (cnt > S.maxc) cnt = S.maxc;
  for (int pass = 0; pass < 2; pass++) {
    for (int s = 0; s < cnt; s++) {
      size_t rb = (size_t)s * 4 * N + e;
      int refA, refB;
      refA = S.s_ref[rb]; refB = S.s_ref[rb + N] & ~(1 << 30);     // written by K4 (contact_refs), bit 30 = sides swapped
      int offA, nA, offB, nB, ns = 0;
      side_dims(S, refA, offA, nA); side_dims(S, refB, offB, nB);
      int o = -1;
      if (nA + nB > 0) {
        if (pass == 0) rs_row_code(refA, refB, nA, nB, ns); else rs_fric_code(refA, refB, nA, nB, ns);
        o = rs_alloc(pos, ns, rs, cap);
        if (o < 0) over = true;
      }
      S.s_ref[rb + (size_t)(2 + pass) * N] = o;
    }
  }
  S.rs_nslots[e] = pos;
  if (over) S.overflow[e] = 1;
}

// K6b, rows part: thread = (dof / fixed-constraint row r, env): write the record K6a reserved
AG_HDN inline void drow_body(int r, int e, const SimDev& S) {
  const int N = S.N;
  const float dt = S.dt;
  int o = S.row_off[(size_t)r * N + e];
  if (o < 0) return;
  float* dst = S.rs_data + ((size_t)e * S.rs_cap + o) * RS_SLOT;
  if (r < 3 * S.ND) {
    int kind = r / S.ND, d = r % S.ND;
    DofRow R;
    if (!dof_row(S, e, kind, d, R)) return;           // cannot happen: K6a saw the same state
    int a = AG_LDG(S.dl_art + d), d0 = AG_LDG(S.art_dl0 + a), nd = AG_LDG(S.art_nd + a), vo = AG_LDG(S.art_voff + a);
    int nd4 = rs_pad4(nd);
    rs_header(dst, RK_ROW_GEN, rs_slots(8 + 2 * nd4), vo, nd4, 0, 0, r);
    dst[4] = R.rhs; dst[5] = R.dinv; dst[6] = R.lo; dst[7] = R.hi;
    for (int i = 0; i < nd4; i++) {
      dst[8 + i] = (i == d - d0) ? R.sgn : 0.f;
      dst[8 + nd4 + i] = i < nd ? R.sgn * S.Minv[((size_t)(d0 + i) * S.ND + d) * N + e] : 0.f;
    }
    return;
  }
  // fixed constraint c, row i: 3 translation + 3 rotation rows
  int c = (r - 3 * S.ND) / 6, i = (r - 3 * S.ND) % 6;
  int refA, refB; bool sw;
  if (!con_sides(S, e, c, refA, refB, sw)) return;
  int ka = AG_LDG(S.con_link + 2 * c), kb = AG_LDG(S.con_link + 2 * c + 1);
  q4 qa = ld4(S.lquat, ka, N, e), qb = ld4(S.lquat, kb, N, e);
  f3 pa = ld3(S.lpos, ka, N, e) + qrot(qa, tv3(S.con_pivot, 2 * c));
  f3 pb = ld3(S.lpos, kb, N, e) + qrot(qb, tv3(S.con_pivot, 2 * c + 1));
  float err;
  if (i < 3) err = comp(pa - pb, i);
  else {
    q4 fa = qmul(qa, tv4(S.con_quat, 2 * c)), fb = qmul(qb, tv4(S.con_quat, 2 * c + 1));
    q4 qe = qmul(fa, qconj(fb));
    if (qe.w < 0.f) qe = q4(-qe.x, -qe.y, -qe.z, -qe.w);
    err = 2.f * comp(f3(qe.x, qe.y, qe.z), i - 3);
  }
  float sg = 1.f;
  if (sw) { f3 tp = pa; pa = pb; pb = tp; sg = -1.f; }     // err was measured before the swap: J is unchanged by it
  float maxi = AG_LDG(S.con_maxforce + c) * dt;
  int offA, nA, offB, nB;
  side_dims(S, refA, offA, nA); side_dims(S, refB, offB, nB);
  int ns, code = rs_row_code(refA, refB, nA, nB, ns);
  int pA = code == RK_ROW_GEN ? rs_pad4(nA) : nA, pB = code == RK_ROW_GEN ? rs_pad4(nB) : nB;
  int oM = code == RK_ROW_F ? 16 : (code == RK_ROW_FF ? 20 : 8 + pA + pB);
  f3 axv(i % 3 == 0 ? 1.f : 0.f, i % 3 == 1 ? 1.f : 0.f, i % 3 == 2 ? 1.f : 0.f);
  f3 lin = i < 3 ? axv : f3(), ang = i < 3 ? f3() : axv;
  float rel = 0.f;
  float diag = emit_side(S, e, refA, pa, lin * sg, ang * sg, dst + 8, dst + oM, pA, rel) +
               emit_side(S, e, refB, pb, lin * (-sg), ang * (-sg), dst + 8 + pA, dst + oM + pA, pB, rel);
  if (!(diag > 1e-20f)) { dst[0] = i2f_bits(RK_PAD | (ns << 4)); return; }
  float dinv = 1.0f / diag;
  rs_header(dst, code, ns, offA, pA, offB, pB, r);
  dst[4] = (-err * S.erp / dt - rel) * dinv; dst[5] = dinv; dst[6] = -maxi; dst[7] = maxi;
}

// K6b: thread = (row, env): rows [0, maxc) are the sorted contacts, rows [maxc, maxc + 3 ND + ngr) the dof and
// fixed-constraint rows
AG_HDN inline void crows_body(int tid, const SimDev& S, const KP&) {
  const int N = S.N;
  int e = tid % N, slot = tid / N;
  if (slot >= S.maxc) { drow_body(slot - S.maxc, e, S); return; }
  int cnt = S.c_count[e]; if (cnt > S.maxc) cnt = S.maxc;
  if (slot >= cnt) return;
  for (int d = 0; d < 3; d++) cf_st(S.s_data, slot, CF_LAM_N + d, N, e, 0.f);
  size_t rb = (size_t)slot * 4 * N + e;
  int refA = S.s_ref[rb], refB = S.s_ref[rb + N], on = S.s_ref[rb + 2 * (size_t)N], of = S.s_ref[rb + 3 * (size_t)N];
  f3 pa(cf_ld(S.s_data, slot, CF_PAX, N, e), cf_ld(S.s_data, slot, CF_PAY, N, e), cf_ld(S.s_data, slot, CF_PAZ, N, e));
  f3 pb(cf_ld(S.s_data, slot, CF_PBX, N, e), cf_ld(S.s_data, slot, CF_PBY, N, e), cf_ld(S.s_data, slot, CF_PBZ, N, e));
  f3 n(cf_ld(S.s_data, slot, CF_NX, N, e), cf_ld(S.s_data, slot, CF_NY, N, e), cf_ld(S.s_data, slot, CF_NZ, N, e));
  float dist = cf_ld(S.s_data, slot, CF_DIST, N, e);
  unsigned pairk = S.s_key[(size_t)slot * N + e] >> 2;
  int ka = AG_LDG(S.col_link + (int)(pairk / (unsigned)S.nc)), kb = AG_LDG(S.col_link + (int)(pairk % (unsigned)S.nc));
  float mu = ld1(S.friction, ka, N, e) * ld1(S.friction, kb, N, e);
  float dt = S.dt;
  // K6a may have swapped the sides (rs_swap_sides); the J entries carry the sign
  float sg = 1.f;
  if (refB & (1 << 30)) { refB &= ~(1 << 30); f3 tp = pa; pa = pb; pb = tp; sg = -1.f; }
  int offA, nA, offB, nB;
  side_dims(S, refA, offA, nA); side_dims(S, refB, offB, nB);
  float* rs = S.rs_data + (size_t)e * S.rs_cap * RS_SLOT;
  const int lam0 = 3 * S.ND + S.ngr + 3 * slot;
  bool live = false;
  if (on >= 0) {
    float* dst = rs + (size_t)on * RS_SLOT;
    int ns, code = rs_row_code(refA, refB, nA, nB, ns);
    int pA = code == RK_ROW_GEN ? rs_pad4(nA) : nA, pB = code == RK_ROW_GEN ? rs_pad4(nB) : nB;
    int oJ = 8, oM = code == RK_ROW_F ? 16 : (code == RK_ROW_FF ? 20 : 8 + pA + pB);
    float rel = 0.f;
    float diag = emit_side(S, e, refA, pa, n * sg, f3(), dst + oJ, dst + oM, pA, rel) +
                 emit_side(S, e, refB, pb, n * (-sg), f3(), dst + oJ + pA, dst + oM + pA, pB, rel);
    if (diag > 1e-20f) {
      float dinv = 1.0f / diag;
      float pen = dist + S.slop;
      float poserr, velerr = -rel;
      if (pen > 0.f) { poserr = 0.f; velerr -= pen / dt; } else poserr = -pen * S.contact_erp / dt;
      rs_header(dst, code, ns, offA, pA, offB, pB, lam0);
      dst[4] = (poserr + velerr) * dinv; dst[5] = dinv; dst[6] = 0.f; dst[7] = 1e30f;
      live = true;
    } else dst[0] = i2f_bits(RK_PAD | (ns << 4));
  }
  if (of >= 0) {
    float* dst = rs + (size_t)of * RS_SLOT;
    int ns, code = rs_fric_code(refA, refB, nA, nB, ns);
    if (!live) { dst[0] = i2f_bits(RK_PAD | (ns << 4)); return; }
    int pA = code == RK_FRIC_GEN ? rs_pad4(nA) : nA, pB = code == RK_FRIC_GEN ? rs_pad4(nB) : nB, P = pA + pB;
    int oJ1, oJ2, oM1, oM2;
    if (code == RK_FRIC_F) { oJ1 = 8; oJ2 = 14; oM1 = 20; oM2 = 26; }
    else if (code == RK_FRIC_FF) { oJ1 = 8; oJ2 = 20; oM1 = 32; oM2 = 44; }
    else { oJ1 = 12; oJ2 = 12 + P; oM1 = 12 + 2 * P; oM2 = 12 + 3 * P; }
    f3 t1, t2; plane_space(n, t1, t2);
    float r1 = 0.f, r2 = 0.f;
    float g1 = emit_side(S, e, refA, pa, t1 * sg, f3(), dst + oJ1, dst + oM1, pA, r1) +
               emit_side(S, e, refB, pb, t1 * (-sg), f3(), dst + oJ1 + pA, dst + oM1 + pA, pB, r1);
    float g2 = emit_side(S, e, refA, pa, t2 * sg, f3(), dst + oJ2, dst + oM2, pA, r2) +
               emit_side(S, e, refB, pb, t2 * (-sg), f3(), dst + oJ2 + pA, dst + oM2 + pA, pB, r2);
    float d1 = g1 > 1e-20f ? 1.0f / g1 : 0.f, d2 = g2 > 1e-20f ? 1.0f / g2 : 0.f;
    rs_header(dst, code, ns, offA, pA, offB, pB, lam0);
    dst[4] = -r1 * d1; dst[5] = d1; dst[6] = -r2 * d2; dst[7] = d2;
    if (code == RK_FRIC_F) dst[2] = mu; else if (code == RK_FRIC_FF) dst[56] = mu; else dst[8] = mu;
  }
}

// ------------------------------------------------------------------ K6c: heaviest-first env order for K7
// The PGS chain of an env is sequential and its length varies 10x between envs (iterations used x
// rows), so CTAs are issued heaviest-first and envs of similar weight share a warp.  Work is predicted
// from this substep's stream length and the previous substep's iteration count.  64-bucket counting
// sort; p.p1 = histogram[64] (zeroed).
AG_HD int pgs_work_bucket(const SimDev& S, int e) {
  int it = S.iters_used[e]; if (it < 1) it = 1;
  int w = it * S.rs_nslots[e];
  int b = 63 - w / 160;                               // heaviest work -> bucket 0
  return b < 0 ? 0 : b;
}
AG_HDN inline void order_hist_body(int e, const SimDev& S, const KP& p) {
  ag_atomic_inc((int*)p.p1 + pgs_work_bucket(S, e));
}
AG_HDN inline void order_scatter_body(int e, const SimDev& S, const KP& p) {
  int pos = ag_atomic_inc((int*)p.p1 + pgs_work_bucket(S, e));
  S.pgs_order[pos] = e;
}
AG_HDN inline void order_prefix_body(int tid, const SimDev&, const KP& p) {
  if (tid != 0) return;
  int* h = (int*)p.p1; int acc = 0;
  for (int b = 0; b < 64; b++) { int c = h[b]; h[b] = acc; acc += c; }
}

// ------------------------------------------------------------------ K7: PGS over the row stream
// The ring is addressed by 32-bit shared-window addresses computed once per lane (`unsigned`), so the
// generic->shared conversion stays out of the row loop.
#if defined(__CUDA_ARCH__)
typedef unsigned rs_addr;
__device__ __forceinline__ rs_addr rs_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rs_bar_init(rs_addr bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
}
// one TMA bulk copy global -> this lane's ring buffer, completion counted in bytes on `bar`
__device__ __forceinline__ void rs_fetch(rs_addr dst, const float*, const float* src, unsigned bytes, rs_addr bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy reads of dst vs. the async write
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void rs_wait(rs_addr bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
#else
typedef size_t rs_addr;
inline rs_addr rs_smem_addr(const void* p) { return (rs_addr)p; }
inline void rs_bar_init(rs_addr) {}
inline void rs_fetch(rs_addr, float* dst, const float* src, unsigned bytes, rs_addr) { memcpy(dst, src, bytes); }
inline void rs_wait(rs_addr, unsigned) {}
#endif

// `sm`: this lane's block of shared memory (rs_lane_floats), `sm_s` its shared-window address, `bar0`: the
// shared-window address of this lane's two mbarriers
AG_HDN inline void pgs_body(int slot, const SimDev& S, const KP&, float* sm, rs_addr sm_s, rs_addr bar0) {
  const int e = S.pgs_order[slot];
  const int N = S.N;
  const int ND = S.ND;
  const int NV = rs_nv(S), NL = rs_nlam(S);
  float* v = sm;
  float* lam = sm + NV;
  float* buf = lam + NL;
#if defined(__CUDA_ARCH__)
  long long t_begin = clock64();
#endif
  const int NBUF = S.rs_nbuf;                     // 2 or 4
  const rs_addr buf_s = sm_s + (rs_addr)(NV + NL) * 4;
  for (int j = 0; j < NBUF; j++) rs_bar_init(bar0 + 8 * j);
#if defined(__CUDA_ARCH__)
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
  const float* rs = S.rs_data + (size_t)e * S.rs_cap * RS_SLOT;
  const int total = S.rs_nslots[e];
  const int nch = (total + RS_CHUNK - 1) / RS_CHUNK;
  const bool resident = nch <= NBUF;              // the whole stream fits the ring: fetch once
  const int gtot = resident ? nch : S.iters * nch; // chunk visits of a full solve
  unsigned pend = 0, phase = 0;                   // per ring buffer: copy in flight / mbarrier phase
  int gi = 0, gck = 0;                            // next chunk visit to issue, and its chunk index
#define RS_ISSUE() do { int j_ = gi & (NBUF - 1); int n_ = total - gck * RS_CHUNK; if (n_ > RS_CHUNK) n_ = RS_CHUNK; \
    rs_fetch(buf_s + j_ * RS_CHUNK * RS_SLOT * 4, buf + j_ * RS_CHUNK * RS_SLOT, rs + (size_t)gck * RS_CHUNK * RS_SLOT, (unsigned)n_ * RS_SLOT * 4, bar0 + 8 * j_); \
    pend |= 1u << j_; gi++; if (++gck == nch) gck = 0; } while (0)
  while (gi < gtot && gi < (resident ? NBUF : NBUF - 1)) RS_ISSUE();
  {
    v4 z; z.x = z.y = z.z = z.w = 0.f;
    for (int i = 0; i < NV; i += 4) stv4(v + i, z);
    for (int i = 0; i < NL; i += 4) stv4(lam + i, z);
  }
  int used = 0;
  int g = 0;                                      // chunk counter (ring position)
  const bool cone = S.cone != 0;
  const int NDp = S.NDp;
  // Register-resident velocity blocks: one free body (side A of F / FF records, side B of generic records) and
  // one articulation (side A of generic records), kept while consecutive records act on the same body.
  int curA = -1;
  v4 a0; v2 a1;
  a0.x = a0.y = a0.z = a0.w = 0.f; a1.x = a1.y = 0.f;
  int curG = -1, nG = 0;
  v4 g0 = a0, g1 = a0, g2 = a0, g3 = a0;
#define RS_FLUSH_A() do { if (curA >= 0) { stv4(v + curA, a0); stv2(v + curA + 4, a1); curA = -1; } } while (0)
#define RS_BIND_A(off) do { if (curA != (off)) { if (curA >= 0) { stv4(v + curA, a0); stv2(v + curA + 4, a1); } curA = (off); a0 = ldv4(v + curA); a1 = ldv2(v + curA + 4); } } while (0)
#define RS_FLUSH_G() do { if (curG >= 0) { stv4(v + curG, g0); if (nG > 4) stv4(v + curG + 4, g1); if (nG > 8) stv4(v + curG + 8, g2); if (nG > 12) stv4(v + curG + 12, g3); curG = -1; } } while (0)
#define RS_BIND_G(off, n) do { if (curG != (off)) { RS_FLUSH_G(); curG = (off); nG = (n); g0 = ldv4(v + curG); if (nG > 4) g1 = ldv4(v + curG + 4); if (nG > 8) g2 = ldv4(v + curG + 8); if (nG > 12) g3 = ldv4(v + curG + 12); } } while (0)
#define RS_AXPY4(X_, M_, S_) do { (X_).x += (M_).x * (S_); (X_).y += (M_).y * (S_); (X_).z += (M_).z * (S_); (X_).w += (M_).w * (S_); } while (0)
#define RS_AXPY4B(X_, M_, S_, N_, T_) do { (X_).x += (M_).x * (S_) + (N_).x * (T_); (X_).y += (M_).y * (S_) + (N_).y * (T_); (X_).z += (M_).z * (S_) + (N_).z * (T_); (X_).w += (M_).w * (S_) + (N_).w * (T_); } while (0)
  for (int it = 0; it < S.iters && nch > 0; it++) {
    float resid = 0.f;
    used = it + 1;
    for (int k = 0; k < nch; k++, g++) {
      const int b = resident ? k : (g & (NBUF - 1));
      if (!resident && gi < gtot) RS_ISSUE();     // keep NBUF-1 chunks in flight ahead of the one being solved
      if ((pend >> b) & 1u) { rs_wait(bar0 + 8 * b, (phase >> b) & 1u); phase ^= 1u << b; pend &= ~(1u << b); }
      const float* cb = buf + b * RS_CHUNK * RS_SLOT;
      int ns = total - k * RS_CHUNK; if (ns > RS_CHUNK) ns = RS_CHUNK;
      v4 hnext = ldv4(cb);                        // header of the chunk's first record
      for (int sl = 0; sl < ns;) {
        const float* r = cb + sl * RS_SLOT;
        // the header was loaded while the previous record was being solved; the rest of the slot comes in one round
        // of vector loads and is consumed from registers
        const v4 h = hnext;
        const v4 c = ldv4(r + 4), q2 = ldv4(r + 8), q3 = ldv4(r + 12), q4v = ldv4(r + 16), q5 = ldv4(r + 20), q6 = ldv4(r + 24), q7 = ldv4(r + 28);
        const int hk = f2i_bits(h.x);
        const int code = hk & 15;
        sl += (hk >> 4) > 0 ? (hk >> 4) : 1;
        if (code >= 2 && sl < ns) hnext = ldv4(cb + sl * RS_SLOT);      // (the F-run loops below fetch their own successors)
        const int wa = f2i_bits(h.y), li = f2i_bits(h.w);
        const int offA = wa & 0xffff;
        if (code < 2) {
          if (code == RK_ROW_F) {
            // Run of rows on one free body: the body's velocity stays in registers and the next record is loaded
            // while the current one is solved, so the dependent chain per row is dot -> clamp -> axpy only.
            RS_BIND_A(offA);
            v4 cc = c, j0 = q2, j1 = q3, m0 = q4v, m1 = q5;
            int lic = li;
            for (;;) {
              const bool more = sl < ns;
              v4 hn = h, cn = c, j0n = q2, j1n = q3, m0n = q4v, m1n = q5;
              if (more) { const float* rn = cb + sl * RS_SLOT; hn = ldv4(rn); cn = ldv4(rn + 4); j0n = ldv4(rn + 8); j1n = ldv4(rn + 12); m0n = ldv4(rn + 16); m1n = ldv4(rn + 20); }
              float jv = (j0.x * a0.x + j0.y * a0.y + j0.z * a0.z) + (j0.w * a0.w + j1.x * a1.x + j1.y * a1.y);
              float l0 = lam[lic];
              float dl = cc.x - jv * cc.y;
              float sum = l0 + dl;
              if (sum < cc.z) { dl = cc.z - l0; sum = cc.z; } else if (sum > cc.w) { dl = cc.w - l0; sum = cc.w; }
              lam[lic] = sum;
              a0.x += m0.x * dl; a0.y += m0.y * dl; a0.z += m0.z * dl; a0.w += m0.w * dl; a1.x += m1.x * dl; a1.y += m1.y * dl;
              resid = fmaxf(resid, dl * dl);
              if (!more || f2i_bits(hn.x) != (RK_ROW_F | (1 << 4)) || (f2i_bits(hn.y) & 0xffff) != curA) { hnext = hn; break; }
              sl += 1; cc = cn; j0 = j0n; j1 = j1n; m0 = m0n; m1 = m1n; lic = f2i_bits(hn.w);
            }
          } else {                                          // RK_FRIC_F
            RS_BIND_A(offA);
            v4 hh = h, cc = c, p2 = q2, p3 = q3, p4 = q4v, p5 = q5, p6 = q6, p7 = q7;
            for (;;) {
              const bool more = sl < ns;
              v4 hn = h, cn = c, n2 = q2, n3 = q3, n4 = q4v, n5 = q5, n6 = q6, n7 = q7;
              if (more) { const float* rn = cb + sl * RS_SLOT; hn = ldv4(rn); cn = ldv4(rn + 4); n2 = ldv4(rn + 8); n3 = ldv4(rn + 12); n4 = ldv4(rn + 16); n5 = ldv4(rn + 20); n6 = ldv4(rn + 24); n7 = ldv4(rn + 28); }
              const int lic = f2i_bits(hh.w);
              float l1 = lam[lic + 1], l2 = lam[lic + 2];
              float lim = hh.z * lam[lic];
              if (!(lim <= 0.f && l1 == 0.f && l2 == 0.f)) {
                float jv1 = (p2.x * a0.x + p2.y * a0.y + p2.z * a0.z) + (p2.w * a0.w + p3.x * a1.x + p3.y * a1.y);
                float jv2 = (p3.z * a0.x + p3.w * a0.y + p4.x * a0.z) + (p4.y * a0.w + p4.z * a1.x + p4.w * a1.y);
                float s1 = l1 + cc.x - jv1 * cc.y, s2 = l2 + cc.z - jv2 * cc.w;
                if (cone) { float m2 = s1 * s1 + s2 * s2; if (m2 > lim * lim) { float kk = lim / sqrtf(m2); s1 *= kk; s2 *= kk; } }
                else { s1 = clampf(s1, -lim, lim); s2 = clampf(s2, -lim, lim); }
                float d1 = s1 - l1, d2 = s2 - l2;
                lam[lic + 1] = s1; lam[lic + 2] = s2;
                a0.x += p5.x * d1 + p6.z * d2; a0.y += p5.y * d1 + p6.w * d2; a0.z += p5.z * d1 + p7.x * d2; a0.w += p5.w * d1 + p7.y * d2;
                a1.x += p6.x * d1 + p7.z * d2; a1.y += p6.y * d1 + p7.w * d2;
                resid = fmaxf(resid, fmaxf(d1 * d1, d2 * d2));
              }
              if (!more || f2i_bits(hn.x) != (RK_FRIC_F | (1 << 4)) || (f2i_bits(hn.y) & 0xffff) != curA) { hnext = hn; break; }
              sl += 1; hh = hn; cc = cn; p2 = n2; p3 = n3; p4 = n4; p5 = n5; p6 = n6; p7 = n7;
            }
          }
        } else if (code < 4) {
          RS_BIND_A(offA);
          float* vb = v + (f2i_bits(h.z) & 0xffff);
          v4 b0 = ldv4(vb); v2 b1 = ldv2(vb + 4);
          if (code == RK_ROW_FF) {
            float jv = (q2.x * a0.x + q2.y * a0.y + q2.z * a0.z) + (q2.w * a0.w + q3.x * a1.x + q3.y * a1.y) +
                       (q3.z * b0.x + q3.w * b0.y + q4v.x * b0.z) + (q4v.y * b0.w + q4v.z * b1.x + q4v.w * b1.y);
            float l0 = lam[li];
            float dl = c.x - jv * c.y;
            float sum = l0 + dl;
            if (sum < c.z) { dl = c.z - l0; sum = c.z; } else if (sum > c.w) { dl = c.w - l0; sum = c.w; }
            lam[li] = sum;
            a0.x += q5.x * dl; a0.y += q5.y * dl; a0.z += q5.z * dl; a0.w += q5.w * dl; a1.x += q6.x * dl; a1.y += q6.y * dl;
            b0.x += q6.z * dl; b0.y += q6.w * dl; b0.z += q7.x * dl; b0.w += q7.y * dl; b1.x += q7.z * dl; b1.y += q7.w * dl;
            stv4(vb, b0); stv2(vb + 4, b1);
            resid = fmaxf(resid, dl * dl);
          } else {                                          // RK_FRIC_FF
            float l1 = lam[li + 1], l2 = lam[li + 2];
            float lim = r[56] * lam[li];
            if (lim <= 0.f && l1 == 0.f && l2 == 0.f) continue;
            float jv1 = (q2.x * a0.x + q2.y * a0.y + q2.z * a0.z) + (q2.w * a0.w + q3.x * a1.x + q3.y * a1.y) +
                        (q3.z * b0.x + q3.w * b0.y + q4v.x * b0.z) + (q4v.y * b0.w + q4v.z * b1.x + q4v.w * b1.y);
            float jv2 = (q5.x * a0.x + q5.y * a0.y + q5.z * a0.z) + (q5.w * a0.w + q6.x * a1.x + q6.y * a1.y) +
                        (q6.z * b0.x + q6.w * b0.y + q7.x * b0.z) + (q7.y * b0.w + q7.z * b1.x + q7.w * b1.y);
            float s1 = l1 + c.x - jv1 * c.y, s2 = l2 + c.z - jv2 * c.w;
            if (cone) { float m2 = s1 * s1 + s2 * s2; if (m2 > lim * lim) { float kk = lim / sqrtf(m2); s1 *= kk; s2 *= kk; } }
            else { s1 = clampf(s1, -lim, lim); s2 = clampf(s2, -lim, lim); }
            float d1 = s1 - l1, d2 = s2 - l2;
            lam[li + 1] = s1; lam[li + 2] = s2;
            const v4 m0 = ldv4(r + 32), m1 = ldv4(r + 36), m2 = ldv4(r + 40), n0 = ldv4(r + 44), n1 = ldv4(r + 48), n2 = ldv4(r + 52);
            a0.x += m0.x * d1 + n0.x * d2; a0.y += m0.y * d1 + n0.y * d2; a0.z += m0.z * d1 + n0.z * d2; a0.w += m0.w * d1 + n0.w * d2;
            a1.x += m1.x * d1 + n1.x * d2; a1.y += m1.y * d1 + n1.y * d2;
            b0.x += m1.z * d1 + n1.z * d2; b0.y += m1.w * d1 + n1.w * d2; b0.z += m2.x * d1 + n2.x * d2; b0.w += m2.y * d1 + n2.y * d2;
            b1.x += m2.z * d1 + n2.z * d2; b1.y += m2.w * d1 + n2.w * d2;
            stv4(vb, b0); stv2(vb + 4, b1);
            resid = fmaxf(resid, fmaxf(d1 * d1, d2 * d2));
          }
        } else if (code < 6) {
          // generic records: side A is an articulation (register block g0..g3), side B nothing, a free body (the
          // a0/a1 block) or a second articulation (shared memory)
          const int wb = f2i_bits(h.z);
          const int nA = wa >> 16, offB = wb & 0xffff, nB = wb >> 16, P = nA + nB;
          const bool bfree = nB == 8 && offB >= NDp;
          RS_BIND_G(offA, nA);
          if (bfree) RS_BIND_A(offB);
          if (code == RK_ROW_GEN) {
            const float* J = r + 8; const float* M = r + 8 + P;
            float jv = dot4(q2, g0);
            if (nA > 4) jv += dot4(q3, g1);
            if (nA > 8) jv += dot4(q4v, g2);
            if (nA > 12) jv += dot4(q5, g3);
            if (bfree) { v4 jb0 = ldv4(J + nA); v2 jb1 = ldv2(J + nA + 4); jv += (jb0.x * a0.x + jb0.y * a0.y + jb0.z * a0.z) + (jb0.w * a0.w + jb1.x * a1.x + jb1.y * a1.y); }
            else for (int i = 0; i < nB; i += 4) jv += dot4(ldv4(J + nA + i), ldv4(v + offB + i));
            float l0 = lam[li];
            float dl = c.x - jv * c.y;
            float sum = l0 + dl;
            if (sum < c.z) { dl = c.z - l0; sum = c.z; } else if (sum > c.w) { dl = c.w - l0; sum = c.w; }
            lam[li] = sum;
            { v4 m = ldv4(M); RS_AXPY4(g0, m, dl); }
            if (nA > 4) { v4 m = ldv4(M + 4); RS_AXPY4(g1, m, dl); }
            if (nA > 8) { v4 m = ldv4(M + 8); RS_AXPY4(g2, m, dl); }
            if (nA > 12) { v4 m = ldv4(M + 12); RS_AXPY4(g3, m, dl); }
            if (bfree) { v4 mb0 = ldv4(M + nA); v2 mb1 = ldv2(M + nA + 4); RS_AXPY4(a0, mb0, dl); a1.x += mb1.x * dl; a1.y += mb1.y * dl; }
            else for (int i = 0; i < nB; i += 4) { v4 m = ldv4(M + nA + i), x = ldv4(v + offB + i); RS_AXPY4(x, m, dl); stv4(v + offB + i, x); }
            resid = fmaxf(resid, dl * dl);
          } else {                                          // RK_FRIC_GEN
            float l1 = lam[li + 1], l2 = lam[li + 2];
            float lim = q2.x * lam[li];
            if (lim <= 0.f && l1 == 0.f && l2 == 0.f) continue;
            const float* J1 = r + 12; const float* J2 = J1 + P; const float* M1 = J2 + P; const float* M2 = M1 + P;
            float jv1 = dot4(ldv4(J1), g0), jv2 = dot4(ldv4(J2), g0);
            if (nA > 4) { jv1 += dot4(ldv4(J1 + 4), g1); jv2 += dot4(ldv4(J2 + 4), g1); }
            if (nA > 8) { jv1 += dot4(ldv4(J1 + 8), g2); jv2 += dot4(ldv4(J2 + 8), g2); }
            if (nA > 12) { jv1 += dot4(ldv4(J1 + 12), g3); jv2 += dot4(ldv4(J2 + 12), g3); }
            if (bfree) {
              v4 x0 = ldv4(J1 + nA), y0 = ldv4(J2 + nA); v2 x1 = ldv2(J1 + nA + 4), y1 = ldv2(J2 + nA + 4);
              jv1 += (x0.x * a0.x + x0.y * a0.y + x0.z * a0.z) + (x0.w * a0.w + x1.x * a1.x + x1.y * a1.y);
              jv2 += (y0.x * a0.x + y0.y * a0.y + y0.z * a0.z) + (y0.w * a0.w + y1.x * a1.x + y1.y * a1.y);
            } else for (int i = 0; i < nB; i += 4) { v4 x = ldv4(v + offB + i); jv1 += dot4(ldv4(J1 + nA + i), x); jv2 += dot4(ldv4(J2 + nA + i), x); }
            float s1 = l1 + c.x - jv1 * c.y, s2 = l2 + c.z - jv2 * c.w;
            if (cone) { float m2 = s1 * s1 + s2 * s2; if (m2 > lim * lim) { float kk = lim / sqrtf(m2); s1 *= kk; s2 *= kk; } }
            else { s1 = clampf(s1, -lim, lim); s2 = clampf(s2, -lim, lim); }
            float d1 = s1 - l1, d2 = s2 - l2;
            lam[li + 1] = s1; lam[li + 2] = s2;
            { v4 m = ldv4(M1), n = ldv4(M2); RS_AXPY4B(g0, m, d1, n, d2); }
            if (nA > 4) { v4 m = ldv4(M1 + 4), n = ldv4(M2 + 4); RS_AXPY4B(g1, m, d1, n, d2); }
            if (nA > 8) { v4 m = ldv4(M1 + 8), n = ldv4(M2 + 8); RS_AXPY4B(g2, m, d1, n, d2); }
            if (nA > 12) { v4 m = ldv4(M1 + 12), n = ldv4(M2 + 12); RS_AXPY4B(g3, m, d1, n, d2); }
            if (bfree) {
              v4 m = ldv4(M1 + nA), n = ldv4(M2 + nA); v2 m1 = ldv2(M1 + nA + 4), n1 = ldv2(M2 + nA + 4);
              RS_AXPY4B(a0, m, d1, n, d2); a1.x += m1.x * d1 + n1.x * d2; a1.y += m1.y * d1 + n1.y * d2;
            } else for (int i = 0; i < nB; i += 4) { v4 m = ldv4(M1 + nA + i), n = ldv4(M2 + nA + i), x = ldv4(v + offB + i); RS_AXPY4B(x, m, d1, n, d2); stv4(v + offB + i, x); }
            resid = fmaxf(resid, fmaxf(d1 * d1, d2 * d2));
          }
        }                                                   // else RK_PAD
      }
    }
    if (S.resid_thr > 0.f && resid <= S.resid_thr) break;
  }
  RS_FLUSH_A();
  RS_FLUSH_G();
#undef RS_FLUSH_A
#undef RS_BIND_A
#undef RS_FLUSH_G
#undef RS_BIND_G
#undef RS_AXPY4
#undef RS_AXPY4B
  // a chunk prefetched for an iteration that never ran must land before the CTA may retire
  for (int j = 0; j < NBUF; j++) if ((pend >> j) & 1u) rs_wait(bar0 + 8 * j, (phase >> j) & 1u);
#undef RS_ISSUE
  // ---- write back
  S.iters_used[e] = used;
#if defined(__CUDA_ARCH__)
  S.pgs_cycles[e] = (int)(clock64() - t_begin);
#endif
  for (int a = 0; a < S.nart; a++) {
    int d0 = AG_LDG(S.art_dl0 + a), nd = AG_LDG(S.art_nd + a), vo = AG_LDG(S.art_voff + a);
    for (int i = 0; i < nd; i++) S.dv[(size_t)(d0 + i) * N + e] = v[vo + i];
  }
  for (int f = 0; f < S.nf; f++)
    for (int c = 0; c < 6; c++) S.dv[(size_t)(ND + 6 * f + c) * N + e] = v[S.NDp + 8 * f + c];
  for (int r = 0; r < 3 * ND; r++) S.dr_lam[(size_t)r * N + e] = lam[r];
  for (int r = 0; r < S.ngr; r++) S.gr_lam[(size_t)r * N + e] = lam[3 * ND + r];
  int cnt = S.c_count[e]; if (cnt > S.maxc) cnt = S.maxc;
  for (int s = 0; s < cnt; s++) {
    const float* l = lam + 3 * ND + S.ngr + 3 * s;
    cf_st(S.s_data, s, CF_LAM_N, N, e, l[0]);
    cf_st(S.s_data, s, CF_LAM_T1, N, e, l[1]);
    cf_st(S.s_data, s, CF_LAM_T2, N, e, l[2]);
  }
}
