// ag_types.h — device-side data layout of the batched simulation.
//
// Layout rule: every per-env quantity is SoA with the env index fastest,
// `[item][component][N]`, so that the 32 lanes of a warp (= 32 consecutive envs) touch one 128 B
// line per scalar.  Template (scene) tables are shared by all envs and read through the read-only
// path; all lanes of a warp read the same address (broadcast).
#pragma once
#include <stdint.h>
#include <stddef.h>

#define AG_MAXND 16          // max DoFs of ONE articulated body (an env may hold several: Jaco 10 + head chains 4 + 4)
#define AG_MAX_HULL 64       // max core vertices per collider
#define AG_CF 20             // floats per contact record
#define AG_CFR 10            // floats per raw (unsorted) contact record: the first 10 fields

// contact record fields (float index within the [AG_CF] record); the RHS / DINV / MU fields are unused since the
// solver rows moved to the packed row stream (ag_solver.cuh), the layout is kept for the read-back kernels
enum {
  CF_PAX = 0, CF_PAY, CF_PAZ, CF_PBX, CF_PBY, CF_PBZ, CF_NX, CF_NY, CF_NZ, CF_DIST,
  CF_RHS_N, CF_DINV_N, CF_RHS_T1, CF_DINV_T1, CF_RHS_T2, CF_DINV_T2, CF_MU, CF_LAM_N, CF_LAM_T1, CF_LAM_T2
};
// body kinds
enum { BK_STATIC = 0, BK_FREE = 1, BK_ART = 2 };

struct SimDev {
  int N;
  // ---- config
  float dt; int iters; float erp, contact_erp, slop, resid_thr, contact_thr, lin_damp, ang_damp, vmax;
  int cone, gyro, maxc;
  // ---- template sizes
  int nb, nl, nc, npair, ncon, nf, nart, ND, nparts, nmovcol, nmovlink, nalllink, ngr /* fixed-constraint rows = 6 ncon */;
  // ---- template tables (device, read-only)
  const int *body_link0, *body_nlinks, *body_kind, *body_idx;
  const float* body_gravity;
  const int *link_body, *link_parent, *link_jtype, *link_dl, *link_haslimit, *link_col0, *link_ncol;
  const float *link_axis, *link_jpos, *link_jquat, *link_com, *link_iquat, *link_inertia, *link_mass, *link_lower, *link_upper;
  const int *col_link, *col_type, *col_v0, *col_nv, *col_p0, *col_np;
  const float *col_radius, *col_thresh, *link_thresh, *col_center, *col_half, *verts, *planes;
  const float* vertq; const int* col_g0;               // core vertices packed 4 at a time ([x0..3][y0..3][z0..3]) for the GJK support search, first group per collider
  float max_thresh;
  const int* pair_link;
  int nslice; const int* pair_slice;                   // [nslice][4] broadphase work items: link a, link b, first collider of a, colliders of a in this slice (<= 64 collider tests each)
  const int *movcol, *movlink, *allcol, *alllink;
  const int* con_link; const float *con_pivot, *con_quat, *con_maxforce;
  const int* free_body; const float* free_invm;
  const int *art_body, *art_dl0, *art_nd, *art_voff;   // art_voff: start of the articulation's block in the solver's velocity vector
  int NDp;                                             // dof part of that vector (every articulation padded to 8)
  const int *dl_link, *dl_parent, *dl_type, *dl_art, *dl_part0, *dl_nparts;
  const float *dl_mass, *dl_mc, *dl_J, *dl_damping;
  const float *pt_mass, *pt_com, *pt_I;
  // ---- motors (per link, shared by envs) + per-env targets
  int* motor_mode; float *motor_kp, *motor_kd, *motor_maxf;
  int* hard_limit;                                     // [nl] clamp q to limits after integration (Human.enforce_joint_limits)
  float *motor_target, *motor_applied;       // [nl][N]
  float* motor_fscale;                       // [nl][N] per-env scale of motor_maxf (Human.strength, human.py:86), or null
  // ---- per-env state
  float *base_pos, *base_quat, *base_lin, *base_ang;   // [nb][3|4][N]
  float *jq, *jqd;                                     // [nl][N]
  float* friction;                                     // [nl][N]
  int* body_mode;                                      // [nb][N]  0 inactive, 1 normal, 2 frozen
  // ---- derived per-env
  float *lpos, *lquat;                                 // [nl][3|4][N]
  float *cmin, *cmax, *lmin, *lmax;                    // [nc|nl][3][N]
  // ---- contacts
  int* cand_count; unsigned *cand, *cand_s; int maxcand; // narrowphase candidates (collider pairs) [maxcand][N]: arrival order / cost order
  int* c_count;                                        // [N]
  int maxraw;                                          // capacity of the raw (arrival-order) contact buffer = 4 maxc
  unsigned *c_key, *s_key;                             // [maxraw][N] unsorted / [maxc][N] sorted
  float *c_data, *s_data;                              // [maxraw][AG_CFR][N] raw / [maxc][AG_CF][N] sorted
  int *s_ref;                                          // [maxc][4][N]: refA, refB, record (rs_enc) of the normal row, friction record offset / 16
  int* overflow;                                       // [N]
  // ---- solver scratch
  float *fcom, *fIinv;                                 // [nf][3|6][N]
  float *jax, *jor;                                    // [ND][3][N] world joint axis / origin
  float* Minv;                                         // [ND][ND][N]
  float* dv;                                           // [ND + 6 nf][N]
  float* dr_lam;                                       // [3 ND][N] impulses of the dof rows: lower limit, upper limit, motor
  float* gr_lam;                                       // [ngr][N] impulses of the fixed-constraint rows
  int* row_off;                                        // [3 ND + ngr][N] record (rs_enc) of each dof / fixed-constraint row (-1: not live)
  int* row_pair;                                       // [3 ND + ngr][N] the row sharing the record of a first row (-1: none)
  float* rs_data; int rs_cap; int* rs_nfloats;         // packed row stream: [N][rs_cap] floats (ag_solver.cuh), used floats [N]
  int* iters_used;                                     // [N]
  int* pgs_order;                                      // heaviest-first env order for K7
  int* pgs_cycles;                                     // [N] SM cycles spent in k_pgs by each env's lane group (diagnostic)
  int* pgs_trips;                                      // [N] records consumed by the env's warp in k_pgs (diagnostic)
};

// kernel-specific small parameter block
struct KP {
  int n;        // number of threads
  int i0, i1, i2, i3;
  float f0, f1;
  const void* p0; void* p1; void* p2; void* p3; void* p4; void* p5;
};
