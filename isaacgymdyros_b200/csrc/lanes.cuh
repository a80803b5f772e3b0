// Lane-group primitives of the multi-lane physics program (physics_lanes.cuh): LPE consecutive lanes of a warp work
// on ONE env. Lane c of a group holds column c of a 6x6 articulated inertia (= row c: the matrices are symmetric) and
// component c of a distributed spatial vector; small vectors are REPLICATED (every lane of the group holds the same
// value, loaded from the same shared-memory address = a broadcast).
//
// The program is written against this small vocabulary so that the very same source also compiles for the host, where
// a group is emulated as a struct of LPE scalars per variable (tests/native/lane_emu.h defines DYROS_LANE_EMU and the
// types below before including the program). On the device everything here is a float / int / bool and a shuffle.
//
//   real            a register value of one lane (device: float)
//   li, lb          lane-varying int / predicate (device: int, bool)
//   bc(g, x, k)     value of x in lane k of the group, in every lane            (shuffle)
//   sh(g, x, src)   value of x in lane src (a lane-varying index) of the group  (shuffle)
//   ld / st         load / store of a REPLICATED value (same address in every lane of the group)
//   ldl / stl       per-lane gather / predicated scatter (lane-varying index)
//   sel(m, a, b)    m ? a : b per lane
//   any(m)          "does this lane take the branch": per-thread on the device (ordinary divergence; the body is then
//                   executed only by lanes with m true), any-lane on the host (the body masks every side effect with m)
//   uni(m)          a predicate known to be the same in all lanes of the group (derived from replicated values)
#pragma once
#include "phys_math.cuh"

#if !defined(DYROS_LANE_EMU)
namespace dyros {

constexpr int LPE = 8;  // lanes per env
typedef int li;
typedef bool lb;

struct Ln {
  int c;           // lane index inside the group, 0 .. LPE-1
  int shift;       // first lane of the group inside the warp
  unsigned gmask;  // the group's lanes (member mask of its shuffles: groups of one warp may diverge from each other)
};
__device__ __forceinline__ Ln make_ln(int lane) {
  Ln g;
  g.c = lane & (LPE - 1);
  g.shift = lane & ~(LPE - 1);
  g.gmask = ((1u << LPE) - 1u) << g.shift;
  return g;
}
__device__ __forceinline__ li lane_index(const Ln& g) { return g.c; }
__device__ __forceinline__ float bc(const Ln& g, float x, int k) { return __shfl_sync(g.gmask, x, k, LPE); }
__device__ __forceinline__ float sh(const Ln& g, float x, int src) { return __shfl_sync(g.gmask, x, src, LPE); }
// number of lanes below this one (rank) / in the whole group (count) whose predicate is set
__device__ __forceinline__ int grank(const Ln& g, bool m) {
  unsigned b = (__ballot_sync(g.gmask, m) >> g.shift) & ((1u << LPE) - 1u);
  return __popc(b & ((1u << g.c) - 1u));
}
__device__ __forceinline__ int gcount(const Ln& g, bool m) {
  return __popc((__ballot_sync(g.gmask, m) >> g.shift) & ((1u << LPE) - 1u));
}
__device__ __forceinline__ float ld(const float* p) { return *p; }
__device__ __forceinline__ void st(float* p, float x) { *p = x; }
__device__ __forceinline__ int ldi(const int* p) { return *p; }
__device__ __forceinline__ void ld4(const float* p, float& a, float& b, float& c, float& d) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  a = t.x; b = t.y; c = t.z; d = t.w;
}
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ float ldl(const float* p, int i) { return p[i]; }
__device__ __forceinline__ int ldli(const int* p, int i) { return p[i]; }
__device__ __forceinline__ void stl(float* p, int i, float x, bool m) {
  if (m) p[i] = x;
}
__device__ __forceinline__ void stli(int* p, int i, int x, bool m) {
  if (m) p[i] = x;
}
__device__ __forceinline__ float sel(bool m, float a, float b) { return m ? a : b; }
__device__ __forceinline__ int seli(bool m, int a, int b) { return m ? a : b; }
__device__ __forceinline__ bool any(bool m) { return m; }
__device__ __forceinline__ bool uni(bool m) { return m; }
__device__ __forceinline__ int unii(int x) { return x; }
__device__ __forceinline__ float rcp_r(float x) { return __frcp_rn(x); }
__device__ __forceinline__ float sqrt_r(float x) { return sqrtf(x); }
__device__ __forceinline__ float int_bits_as_real(int x) { return __int_as_float(x); }
__device__ __forceinline__ int real_bits_as_int(float x) { return __float_as_int(x); }
__device__ __forceinline__ void lane_fence() { __syncwarp(); }  // orders shared-memory traffic between the lanes of a warp

}  // namespace dyros
#endif
