// Small fixed-size linear algebra for the articulated-body recursions (3-vectors, 3x3 blocks, 6-D spatial
// vectors, symmetric 6x6 articulated inertias). Plain structs of scalars so everything stays in registers.
// Host+device: the same code is compiled by g++ for the CPU emulation used in tests (tests/native/).
#pragma once
#include <math.h>

#if defined(__CUDACC__) && !defined(DYROS_LANE_EMU)
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

namespace dyros {

#ifndef DYROS_REAL
#define DYROS_REAL float
#endif
typedef DYROS_REAL real;

struct V3 {
  real x, y, z;
};
struct M3 {  // row-major
  real a[9];
};
struct S3 {  // symmetric 3x3
  real xx, yy, zz, xy, xz, yz;
};
struct SV {  // spatial motion [w; v] or force [n; f]
  V3 w, v;
};
struct ABI {  // [[I, H], [H^T, M]] acting on [w; v]; also used for inverse inertias [[A, B], [B^T, C]] on [n; f]
  S3 I;
  M3 H;
  S3 M;
};

HD V3 v3(real x, real y, real z) { return V3{x, y, z}; }
HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
HD V3 operator*(real s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
HD V3 neg(V3 a) { return V3{-a.x, -a.y, -a.z}; }
HD real dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
HD V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
HD V3 ld3(const real* p) { return V3{p[0], p[1], p[2]}; }
#ifndef DYROS_REAL_IS_FLOAT_ONLY
template <class F>
HD V3 ld3_f(const F* p) { return V3{(real)p[0], (real)p[1], (real)p[2]}; }  // from float tables / tensors
#endif
#if defined(DYROS_LANE_EMU)  // host emulation of a lane group (tests/native/lane_emu.h): `real` is a vector of lanes
HD real fmin_r(real a, real b) { return sel(a < b, a, b); }
#else
HD real fmin_r(real a, real b) { return a < b ? a : b; }
#endif
HD void sincos_r(real x, real* s, real* c) {
#if defined(__CUDA_ARCH__)
  float sf, cf;
  sincosf((float)x, &sf, &cf);
  *s = sf;
  *c = cf;
#else
  *s = sin(x);
  *c = cos(x);
#endif
}
HD void st3(real* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

HD V3 mul(const M3& m, V3 v) {
  return V3{m.a[0] * v.x + m.a[1] * v.y + m.a[2] * v.z, m.a[3] * v.x + m.a[4] * v.y + m.a[5] * v.z,
            m.a[6] * v.x + m.a[7] * v.y + m.a[8] * v.z};
}
HD V3 mulT(const M3& m, V3 v) {  // m^T v
  return V3{m.a[0] * v.x + m.a[3] * v.y + m.a[6] * v.z, m.a[1] * v.x + m.a[4] * v.y + m.a[7] * v.z,
            m.a[2] * v.x + m.a[5] * v.y + m.a[8] * v.z};
}
HD V3 mul(const S3& s, V3 v) {
  return V3{s.xx * v.x + s.xy * v.y + s.xz * v.z, s.xy * v.x + s.yy * v.y + s.yz * v.z,
            s.xz * v.x + s.yz * v.y + s.zz * v.z};
}
HD M3 mul(const M3& a, const M3& b) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c.a[3 * i + j] = a.a[3 * i] * b.a[j] + a.a[3 * i + 1] * b.a[3 + j] + a.a[3 * i + 2] * b.a[6 + j];
  return c;
}
HD M3 mulABt(const M3& a, const M3& b) {  // a b^T
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c.a[3 * i + j] = a.a[3 * i] * b.a[3 * j] + a.a[3 * i + 1] * b.a[3 * j + 1] + a.a[3 * i + 2] * b.a[3 * j + 2];
  return c;
}
HD M3 mulAtB(const M3& a, const M3& b) {  // a^T b
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c.a[3 * i + j] = a.a[i] * b.a[j] + a.a[3 + i] * b.a[3 + j] + a.a[6 + i] * b.a[6 + j];
  return c;
}
HD M3 transpose(const M3& a) { return M3{{a.a[0], a.a[3], a.a[6], a.a[1], a.a[4], a.a[7], a.a[2], a.a[5], a.a[8]}}; }
HD M3 full(const S3& s) { return M3{{s.xx, s.xy, s.xz, s.xy, s.yy, s.yz, s.xz, s.yz, s.zz}}; }
HD S3 sym_of(const M3& m) { return S3{m.a[0], m.a[4], m.a[8], m.a[1], m.a[2], m.a[5]}; }  // upper triangle
HD S3 operator+(S3 a, S3 b) { return S3{a.xx + b.xx, a.yy + b.yy, a.zz + b.zz, a.xy + b.xy, a.xz + b.xz, a.yz + b.yz}; }
HD M3 operator+(const M3& a, const M3& b) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 9; ++i) c.a[i] = a.a[i] + b.a[i];
  return c;
}
HD M3 operator-(const M3& a, const M3& b) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 9; ++i) c.a[i] = a.a[i] - b.a[i];
  return c;
}
// skew(r) * m  (rows of the product are r x (columns of m))
HD M3 skew_mul(V3 r, const M3& m) {
  M3 c;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    V3 col = cross(r, V3{m.a[j], m.a[3 + j], m.a[6 + j]});
    c.a[j] = col.x; c.a[3 + j] = col.y; c.a[6 + j] = col.z;
  }
  return c;
}
// m * skew(r)  (row i of the product is (row i of m) x r ... with sign: row_i(m) * skew(r) = -(r x row_i)^T = (row_i x r)^T)
HD M3 mul_skew(const M3& m, V3 r) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    V3 row = cross(V3{m.a[3 * i], m.a[3 * i + 1], m.a[3 * i + 2]}, r);
    c.a[3 * i] = row.x; c.a[3 * i + 1] = row.y; c.a[3 * i + 2] = row.z;
  }
  return c;
}
// E S E^T for symmetric S (result symmetric)
HD S3 rot_sym(const M3& E, const S3& s) {
  M3 t = mul(E, full(s));
  S3 r;
  r.xx = t.a[0] * E.a[0] + t.a[1] * E.a[1] + t.a[2] * E.a[2];
  r.yy = t.a[3] * E.a[3] + t.a[4] * E.a[4] + t.a[5] * E.a[5];
  r.zz = t.a[6] * E.a[6] + t.a[7] * E.a[7] + t.a[8] * E.a[8];
  r.xy = t.a[0] * E.a[3] + t.a[1] * E.a[4] + t.a[2] * E.a[5];
  r.xz = t.a[0] * E.a[6] + t.a[1] * E.a[7] + t.a[2] * E.a[8];
  r.yz = t.a[3] * E.a[6] + t.a[4] * E.a[7] + t.a[5] * E.a[8];
  return r;
}
// E^T S E
HD S3 rotT_sym(const M3& E, const S3& s) { return rot_sym(transpose(E), s); }

HD SV operator+(SV a, SV b) { return SV{a.w + b.w, a.v + b.v}; }
HD SV operator-(SV a, SV b) { return SV{a.w - b.w, a.v - b.v}; }
HD SV operator*(real s, SV a) { return SV{s * a.w, s * a.v}; }
HD real dot(SV a, SV b) { return dot(a.w, b.w) + dot(a.v, b.v); }
HD SV ld6(const real* p) { return SV{ld3(p), ld3(p + 3)}; }
HD void st6(real* p, SV a) { st3(p, a.w); st3(p + 3, a.v); }
HD SV sv_zero() { return SV{V3{0, 0, 0}, V3{0, 0, 0}}; }

// Pluecker transforms for X = (E, r): E rotates parent coordinates into child coordinates, r = child origin in
// parent coordinates.
HD SV xform_motion(const M3& E, V3 r, SV m) {  // parent -> child
  return SV{mul(E, m.w), mul(E, m.v - cross(r, m.w))};
}
HD SV xform_force_T(const M3& E, V3 r, SV f) {  // child -> parent (X^T f)
  V3 fp = mulT(E, f.v);
  return SV{mulT(E, f.w) + cross(r, fp), fp};
}
// v x m and v x* f
HD SV crm(SV v, SV m) { return SV{cross(v.w, m.w), cross(v.w, m.v) + cross(v.v, m.w)}; }
HD SV crf(SV v, SV f) { return SV{cross(v.w, f.w) + cross(v.v, f.v), cross(v.w, f.v)}; }

// rigid-body inertia from (m, h = m c, Ibar about the origin)
HD ABI abi_rigid(real m, V3 h, S3 Ibar) {
  ABI a;
  a.I = Ibar;
  a.H = M3{{0, -h.z, h.y, h.z, 0, -h.x, -h.y, h.x, 0}};
  a.M = S3{m, m, m, 0, 0, 0};
  return a;
}
HD SV mul(const ABI& a, SV m) {  // [[I,H],[H^T,M]] [w; v]
  return SV{mul(a.I, m.w) + mul(a.H, m.v), mulT(a.H, m.w) + mul(a.M, m.v)};
}
HD ABI operator+(const ABI& a, const ABI& b) { return ABI{a.I + b.I, a.H + b.H, a.M + b.M}; }
// a - s * u u^T
HD ABI rank1_sub(const ABI& a, SV u, real s) {
  ABI r = a;
  V3 sw = s * u.w, sv = s * u.v;
  r.I.xx -= sw.x * u.w.x; r.I.yy -= sw.y * u.w.y; r.I.zz -= sw.z * u.w.z;
  r.I.xy -= sw.x * u.w.y; r.I.xz -= sw.x * u.w.z; r.I.yz -= sw.y * u.w.z;
  r.M.xx -= sv.x * u.v.x; r.M.yy -= sv.y * u.v.y; r.M.zz -= sv.z * u.v.z;
  r.M.xy -= sv.x * u.v.y; r.M.xz -= sv.x * u.v.z; r.M.yz -= sv.y * u.v.z;
  r.H.a[0] -= sw.x * u.v.x; r.H.a[1] -= sw.x * u.v.y; r.H.a[2] -= sw.x * u.v.z;
  r.H.a[3] -= sw.y * u.v.x; r.H.a[4] -= sw.y * u.v.y; r.H.a[5] -= sw.y * u.v.z;
  r.H.a[6] -= sw.z * u.v.x; r.H.a[7] -= sw.z * u.v.y; r.H.a[8] -= sw.z * u.v.z;
  return r;
}
// Articulated inertia of a child expressed at its parent: X^T Ia X  (force-type congruence)
HD ABI abi_to_parent(const M3& E, V3 r, const ABI& a) {
  M3 Et = transpose(E);
  S3 I1 = rot_sym(Et, a.I);
  S3 M1 = rot_sym(Et, a.M);
  M3 H1 = mul(mul(Et, a.H), E);                 // E^T H E
  M3 RM = skew_mul(r, full(M1));                // r~ M'
  M3 Hp = H1 + RM;                              // H' + r~ M'
  M3 A = skew_mul(r, transpose(H1));            // r~ H'^T
  M3 B = mul_skew(Hp, r);                       // Hp r~
  ABI p;
  p.M = M1;
  p.H = Hp;
  // I' + r~ H'^T - Hp r~  (symmetric by construction; take the symmetric part for round-off)
  p.I.xx = I1.xx + A.a[0] - B.a[0];
  p.I.yy = I1.yy + A.a[4] - B.a[4];
  p.I.zz = I1.zz + A.a[8] - B.a[8];
  p.I.xy = I1.xy + (real)0.5 * ((A.a[1] - B.a[1]) + (A.a[3] - B.a[3]));
  p.I.xz = I1.xz + (real)0.5 * ((A.a[2] - B.a[2]) + (A.a[6] - B.a[6]));
  p.I.yz = I1.yz + (real)0.5 * ((A.a[5] - B.a[5]) + (A.a[7] - B.a[7]));
  return p;
}
// ---- the same transforms for a pure translation (all quantities in world axes, reference point moved by r)
HD SV shift_motion(V3 r, SV m) { return SV{m.w, m.v - cross(r, m.w)}; }   // parent origin -> child origin (r = child - parent)
HD SV shift_force_T(V3 r, SV f) { return SV{f.w + cross(r, f.v), f.v}; }  // child origin -> parent origin
HD ABI abi_shift_to_parent(V3 r, const ABI& a) {
  M3 RM = skew_mul(r, full(a.M));               // r~ M
  M3 Hp = a.H + RM;                             // H + r~ M
  M3 A = skew_mul(r, transpose(a.H));           // r~ H^T
  M3 B = mul_skew(Hp, r);                       // Hp r~
  ABI p;
  p.M = a.M;
  p.H = Hp;
  p.I.xx = a.I.xx + A.a[0] - B.a[0];
  p.I.yy = a.I.yy + A.a[4] - B.a[4];
  p.I.zz = a.I.zz + A.a[8] - B.a[8];
  p.I.xy = a.I.xy + (real)0.5 * ((A.a[1] - B.a[1]) + (A.a[3] - B.a[3]));
  p.I.xz = a.I.xz + (real)0.5 * ((A.a[2] - B.a[2]) + (A.a[6] - B.a[6]));
  p.I.yz = a.I.yz + (real)0.5 * ((A.a[5] - B.a[5]) + (A.a[7] - B.a[7]));
  return p;
}
HD ABI inv_shift_to_child(V3 r, const ABI& o) {
  M3 A = full(o.I);
  M3 Bp = o.H + mul_skew(A, r);
  M3 RB = skew_mul(r, o.H);
  M3 BtR = mul_skew(transpose(Bp), r);
  M3 Cp = full(o.M) - RB + BtR;
  ABI c;
  c.I = o.I;
  c.H = Bp;
  c.M = S3{Cp.a[0], Cp.a[4], Cp.a[8], (real)0.5 * (Cp.a[1] + Cp.a[3]), (real)0.5 * (Cp.a[2] + Cp.a[6]),
           (real)0.5 * (Cp.a[5] + Cp.a[7])};
  return c;
}
// Inverse inertia (maps force to motion) of the parent expressed at the child: X Om X^T (motion-type congruence)
HD ABI inv_to_child(const M3& E, V3 r, const ABI& o) {
  // shift: A' = A; B' = B + A r~; C' = C - r~ B + B'^T r~
  M3 A = full(o.I);
  M3 Bp = o.H + mul_skew(A, r);
  M3 RB = skew_mul(r, o.H);
  M3 BtR = mul_skew(transpose(Bp), r);
  M3 C = full(o.M);
  M3 Cp = C - RB + BtR;
  ABI c;
  c.I = rot_sym(E, o.I);
  c.H = mulABt(mul(E, Bp), E);  // E B' E^T
  S3 Cs = S3{Cp.a[0], Cp.a[4], Cp.a[8], (real)0.5 * (Cp.a[1] + Cp.a[3]), (real)0.5 * (Cp.a[2] + Cp.a[6]),
             (real)0.5 * (Cp.a[5] + Cp.a[7])};
  c.M = rot_sym(E, Cs);
  return c;
}
// a + s * (x y^T + y x^T) style helpers for the inverse-inertia recursion: a -= s y^T + y s^T ; a += k s s^T
// where s = [axis; 0] (revolute). Only the angular rows/columns of s are non-zero.
HD ABI inv_joint_update(const ABI& a, V3 ax, SV y, real k) {
  ABI r = a;
  // [[A, B],[B^T, C]] -= [ax;0] [y.w; y.v]^T + [y.w; y.v] [ax;0]^T ;  += k [ax;0][ax;0]^T
  r.I.xx += k * ax.x * ax.x - 2 * ax.x * y.w.x;
  r.I.yy += k * ax.y * ax.y - 2 * ax.y * y.w.y;
  r.I.zz += k * ax.z * ax.z - 2 * ax.z * y.w.z;
  r.I.xy += k * ax.x * ax.y - ax.x * y.w.y - ax.y * y.w.x;
  r.I.xz += k * ax.x * ax.z - ax.x * y.w.z - ax.z * y.w.x;
  r.I.yz += k * ax.y * ax.z - ax.y * y.w.z - ax.z * y.w.y;
  // B block (rows angular, cols linear): -= ax * y.v^T
  r.H.a[0] -= ax.x * y.v.x; r.H.a[1] -= ax.x * y.v.y; r.H.a[2] -= ax.x * y.v.z;
  r.H.a[3] -= ax.y * y.v.x; r.H.a[4] -= ax.y * y.v.y; r.H.a[5] -= ax.y * y.v.z;
  r.H.a[6] -= ax.z * y.v.x; r.H.a[7] -= ax.z * y.v.y; r.H.a[8] -= ax.z * y.v.z;
  return r;
}

// quaternion xyzw -> rotation (body to world)
HD M3 quat_to_mat(real x, real y, real z, real w) {
  return M3{{1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z),
             2 * (y * z - w * x), 2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)}};
}
// Rodrigues rotation about a unit axis by angle q, transposed (coordinates of a parent-fixed vector in the rotated frame)
HD M3 axis_rot_T(V3 a, real s, real c) {
  real t = 1 - c;
  M3 R{{c + t * a.x * a.x, t * a.x * a.y - s * a.z, t * a.x * a.z + s * a.y, t * a.x * a.y + s * a.z, c + t * a.y * a.y,
        t * a.y * a.z - s * a.x, t * a.x * a.z - s * a.y, t * a.y * a.z + s * a.x, c + t * a.z * a.z}};
  return transpose(R);
}

// Inverse of a symmetric positive definite 3x3 (adjugate / determinant).
HD S3 spd3_inverse(const S3& s) {
  real cxx = s.yy * s.zz - s.yz * s.yz, cyy = s.xx * s.zz - s.xz * s.xz, czz = s.xx * s.yy - s.xy * s.xy;
  real cxy = s.xz * s.yz - s.xy * s.zz, cxz = s.xy * s.yz - s.xz * s.yy, cyz = s.xy * s.xz - s.xx * s.yz;
  real inv = 1 / (s.xx * cxx + s.xy * cxy + s.xz * cxz);
  return S3{inv * cxx, inv * cyy, inv * czz, inv * cxy, inv * cxz, inv * cyz};
}
// Inverse of a symmetric positive definite [[I, H], [H^T, M]] through the Schur complement of M (M: the
// articulated mass block, I - H M^-1 H^T: the rotational inertia about the centre of mass; both well conditioned):
// [[S^-1, -S^-1 P], [-P^T S^-1, M^-1 + P^T S^-1 P]] with P = H M^-1, S = I - P H^T.
HD ABI abi_inverse_spd(const ABI& a) {
  const M3 Mi = full(spd3_inverse(a.M));
  const M3 P = mul(a.H, Mi);
  const M3 PHt = mulABt(P, a.H);
  const S3 S{a.I.xx - PHt.a[0], a.I.yy - PHt.a[4], a.I.zz - PHt.a[8],
             a.I.xy - (real)0.5 * (PHt.a[1] + PHt.a[3]), a.I.xz - (real)0.5 * (PHt.a[2] + PHt.a[6]),
             a.I.yz - (real)0.5 * (PHt.a[5] + PHt.a[7])};
  const M3 Si = full(spd3_inverse(S));
  const M3 Q = mul(Si, P);  // S^-1 P
  const M3 PtQ = mulAtB(P, Q);
  ABI o;
  o.I = sym_of(Si);
  o.H = M3{{-Q.a[0], -Q.a[1], -Q.a[2], -Q.a[3], -Q.a[4], -Q.a[5], -Q.a[6], -Q.a[7], -Q.a[8]}};
  o.M = S3{Mi.a[0] + PtQ.a[0], Mi.a[4] + PtQ.a[4], Mi.a[8] + PtQ.a[8], Mi.a[1] + (real)0.5 * (PtQ.a[1] + PtQ.a[3]),
           Mi.a[2] + (real)0.5 * (PtQ.a[2] + PtQ.a[6]), Mi.a[5] + (real)0.5 * (PtQ.a[5] + PtQ.a[7])};
  return o;
}

// In-place inverse of a symmetric positive definite 6x6 (row-major full storage) by Cholesky.
HD void spd6_inverse(real* a) {
  real L[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) L[i] = 0;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    real d = a[7 * j];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k < j) d -= L[6 * j + k] * L[6 * j + k];
    real inv = 1 / sqrt(d);
    L[7 * j] = d * inv;
#pragma unroll
    for (int i = 0; i < 6; ++i)
      if (i > j) {
        real s = a[6 * i + j];
#pragma unroll
        for (int k = 0; k < 6; ++k)
          if (k < j) s -= L[6 * i + k] * L[6 * j + k];
        L[6 * i + j] = s * inv;
      }
  }
  // Linv (lower triangular)
  real Li[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) Li[i] = 0;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    Li[7 * j] = 1 / L[7 * j];
#pragma unroll
    for (int i = 0; i < 6; ++i)
      if (i > j) {
        real s = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k)
          if (k >= j && k < i) s -= L[6 * i + k] * Li[6 * k + j];
        Li[6 * i + j] = s / L[7 * i];
      }
  }
  // A^-1 = Linv^T Linv
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      real s = 0;
#pragma unroll
      for (int k = 0; k < 6; ++k)
        if (k >= i && k >= j) s += Li[6 * k + i] * Li[6 * k + j];
      a[6 * i + j] = s;
    }
}
HD void abi_to_full(const ABI& a, real* m) {  // 6x6 row-major
  M3 I = full(a.I), M = full(a.M);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      m[6 * i + j] = I.a[3 * i + j];
      m[6 * i + 3 + j] = a.H.a[3 * i + j];
      m[6 * (3 + i) + j] = a.H.a[3 * j + i];
      m[6 * (3 + i) + 3 + j] = M.a[3 * i + j];
    }
}
HD ABI abi_from_full(const real* m) {
  ABI a;
  a.I = S3{m[0], m[7], m[14], m[1], m[2], m[8]};
  a.M = S3{m[21], m[28], m[35], m[22], m[23], m[29]};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) a.H.a[3 * i + j] = m[6 * i + 3 + j];
  return a;
}
HD void st_abi(real* p, const ABI& a) {
  p[0] = a.I.xx; p[1] = a.I.yy; p[2] = a.I.zz; p[3] = a.I.xy; p[4] = a.I.xz; p[5] = a.I.yz;
#pragma unroll
  for (int i = 0; i < 9; ++i) p[6 + i] = a.H.a[i];
  p[15] = a.M.xx; p[16] = a.M.yy; p[17] = a.M.zz; p[18] = a.M.xy; p[19] = a.M.xz; p[20] = a.M.yz;
}
HD ABI ld_abi(const real* p) {
  ABI a;
  a.I = S3{p[0], p[1], p[2], p[3], p[4], p[5]};
#pragma unroll
  for (int i = 0; i < 9; ++i) a.H.a[i] = p[6 + i];
  a.M = S3{p[15], p[16], p[17], p[18], p[19], p[20]};
  return a;
}
template <class F>
HD M3 ld_m3_f(const F* p) {
  M3 m;
#pragma unroll
  for (int i = 0; i < 9; ++i) m.a[i] = (real)p[i];
  return m;
}
HD M3 ld_m3(const real* p) {
  M3 m;
#pragma unroll
  for (int i = 0; i < 9; ++i) m.a[i] = p[i];
  return m;
}
HD void st_m3(real* p, const M3& m) {
#pragma unroll
  for (int i = 0; i < 9; ++i) p[i] = m.a[i];
}

}  // namespace dyros
