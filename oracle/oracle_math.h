// oracle_math.h — small vector / quaternion / spatial algebra for the CPU oracle.
// TEST INFRASTRUCTURE ONLY (see oracle/README.md): nothing under assistive_gym_b200/ may use this.
#pragma once
#include <cmath>
#include <cstring>
#include <algorithm>

#ifdef ORACLE_FLOAT
typedef float real;
#else
typedef double real;
#endif

struct V3 {
  real x, y, z;
  V3() : x(0), y(0), z(0) {}
  V3(real a, real b, real c) : x(a), y(b), z(c) {}
  real& operator[](int i) { return (&x)[i]; }
  real operator[](int i) const { return (&x)[i]; }
};
static inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
static inline V3 operator*(V3 a, real s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator*(real s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }
static inline V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
static inline V3& operator-=(V3& a, V3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
static inline real dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline real norm(V3 a) { return std::sqrt(dot(a, a)); }

struct Quat { real x, y, z, w; Quat() : x(0), y(0), z(0), w(1) {} Quat(real a, real b, real c, real d) : x(a), y(b), z(c), w(d) {} };
static inline Quat qmul(Quat a, Quat b) {
  return Quat(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
              a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w,
              a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
static inline Quat qconj(Quat q) { return Quat(-q.x, -q.y, -q.z, q.w); }
static inline Quat qnormalize(Quat q) {
  real n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return Quat(q.x / n, q.y / n, q.z / n, q.w / n);
}
static inline V3 qrot(Quat q, V3 v) {
  V3 u(q.x, q.y, q.z);
  V3 t = cross(u, v) * real(2);
  return v + t * q.w + cross(u, t);
}
static inline Quat qaxis(V3 axis, real angle) {
  real s = std::sin(angle / 2);
  return Quat(axis.x * s, axis.y * s, axis.z * s, std::cos(angle / 2));
}
// exponential map: rotation by vector w (angle = |w|)
static inline Quat qexp(V3 w) {
  real a = norm(w);
  if (a < real(1e-12)) return qnormalize(Quat(w.x / 2, w.y / 2, w.z / 2, 1));
  real s = std::sin(a / 2) / a;
  return Quat(w.x * s, w.y * s, w.z * s, std::cos(a / 2));
}

struct M3 {
  real m[3][3];
  M3() { std::memset(m, 0, sizeof(m)); }
  static M3 ident() { M3 r; r.m[0][0] = r.m[1][1] = r.m[2][2] = 1; return r; }
  static M3 diag(V3 d) { M3 r; r.m[0][0] = d.x; r.m[1][1] = d.y; r.m[2][2] = d.z; return r; }
};
static inline M3 qmat(Quat q) {
  M3 r;
  real x = q.x, y = q.y, z = q.z, w = q.w;
  r.m[0][0] = 1 - 2 * (y * y + z * z); r.m[0][1] = 2 * (x * y - z * w); r.m[0][2] = 2 * (x * z + y * w);
  r.m[1][0] = 2 * (x * y + z * w); r.m[1][1] = 1 - 2 * (x * x + z * z); r.m[1][2] = 2 * (y * z - x * w);
  r.m[2][0] = 2 * (x * z - y * w); r.m[2][1] = 2 * (y * z + x * w); r.m[2][2] = 1 - 2 * (x * x + y * y);
  return r;
}
static inline V3 operator*(const M3& a, V3 v) {
  return V3(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
            a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z);
}
static inline M3 operator*(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { real s = 0; for (int k = 0; k < 3; k++) s += a.m[i][k] * b.m[k][j]; r.m[i][j] = s; }
  return r;
}
static inline M3 transpose(const M3& a) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i]; return r; }
static inline M3 operator+(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
static inline M3 skew(V3 v) { M3 r; r.m[0][1] = -v.z; r.m[0][2] = v.y; r.m[1][0] = v.z; r.m[1][2] = -v.x; r.m[2][0] = -v.y; r.m[2][1] = v.x; return r; }
static inline M3 inverse(const M3& a) {
  M3 r;
  real det = a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) - a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0]) + a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
  real id = 1 / det;
  r.m[0][0] = (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) * id;
  r.m[0][1] = (a.m[0][2] * a.m[2][1] - a.m[0][1] * a.m[2][2]) * id;
  r.m[0][2] = (a.m[0][1] * a.m[1][2] - a.m[0][2] * a.m[1][1]) * id;
  r.m[1][0] = (a.m[1][2] * a.m[2][0] - a.m[1][0] * a.m[2][2]) * id;
  r.m[1][1] = (a.m[0][0] * a.m[2][2] - a.m[0][2] * a.m[2][0]) * id;
  r.m[1][2] = (a.m[0][2] * a.m[1][0] - a.m[0][0] * a.m[1][2]) * id;
  r.m[2][0] = (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]) * id;
  r.m[2][1] = (a.m[0][1] * a.m[2][0] - a.m[0][0] * a.m[2][1]) * id;
  r.m[2][2] = (a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0]) * id;
  return r;
}

// Spatial vectors, world frame, referred to the world origin.  Motion: (w, v_O); force: (n_O, f).
struct SV {
  V3 a, l;  // angular part, linear part
  SV() {}
  SV(V3 a_, V3 l_) : a(a_), l(l_) {}
};
static inline SV operator+(SV p, SV q) { return SV(p.a + q.a, p.l + q.l); }
static inline SV operator-(SV p, SV q) { return SV(p.a - q.a, p.l - q.l); }
static inline SV operator*(SV p, real s) { return SV(p.a * s, p.l * s); }
static inline real sdot(SV m, SV f) { return dot(m.a, f.a) + dot(m.l, f.l); }
// motion x motion
static inline SV crm(SV v, SV m) { return SV(cross(v.a, m.a), cross(v.a, m.l) + cross(v.l, m.a)); }
// motion x* force
static inline SV crf(SV v, SV f) { return SV(cross(v.a, f.a) + cross(v.l, f.l), cross(v.a, f.l)); }

// 6x6 spatial inertia (symmetric, maps motion -> force), stored as 4 3x3 blocks [[A, B],[B^T, D]]
struct SI {
  M3 A, B, D;  // A: ang-ang, B: ang-lin, D: lin-lin
};
static inline SI operator+(const SI& p, const SI& q) { SI r; r.A = p.A + q.A; r.B = p.B + q.B; r.D = p.D + q.D; return r; }
static inline SV operator*(const SI& I, SV v) { return SV(I.A * v.a + I.B * v.l, transpose(I.B) * v.a + I.D * v.l); }
// rigid body inertia about the world origin: mass m, COM c (world), rotational inertia Ic (world axes, about COM)
static inline SI rigid_inertia(real m, V3 c, const M3& Ic) {
  SI r;
  M3 cx = skew(c);
  M3 cxT = transpose(cx);
  M3 t = cx * cxT;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    r.A.m[i][j] = Ic.m[i][j] + m * t.m[i][j];
    r.B.m[i][j] = m * cx.m[i][j];
    r.D.m[i][j] = (i == j) ? m : 0;
  }
  return r;
}
// I - U U^T / d
static inline SI sub_outer(const SI& I, SV U, real invd) {
  SI r = I;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    r.A.m[i][j] -= U.a[i] * U.a[j] * invd;
    r.B.m[i][j] -= U.a[i] * U.l[j] * invd;
    r.D.m[i][j] -= U.l[i] * U.l[j] * invd;
  }
  return r;
}
