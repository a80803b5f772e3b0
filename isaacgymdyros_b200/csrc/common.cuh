// Shared device helpers: error plumbing, Philox4x32-10 counter RNG, warp utilities.
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dyros_b200.h"

namespace dyros {

void set_error(const char* fmt, ...);
#define DY_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      dyros::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)
#define DY_LAUNCH_CHECK() DY_CUDA(cudaGetLastError())

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// The four kernels of one env-step are launched back to back on one stream. With the programmatic-stream-serialisation
// attribute a kernel's CTAs may become resident (and run their preamble, e.g. staging the model tables) while the
// previous kernel drains; griddepcontrol.wait then blocks until the previous grid has completed and flushed.
// Without the attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool no_pdl = getenv("DYROS_NO_PDL") != nullptr;  // (diagnostics: every launch fully serialised)
  cfg.numAttrs = pdl && !no_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// Every kernel of the library asks for the same L1 / shared-memory split as the physics kernels, which need all of the
// shared memory of an SM (226 KB). Measured on B200 (profiles/r2_carveout.md): a stream that alternates between kernels
// with different preferred carve-outs pays ~25 us per switch (the SMs drain and reconfigure), which was ~50 us of a
// ~147 us fused step; with one carve-out everywhere the same launches take the sum of their kernels (~96 us).
template <class K>
inline cudaError_t prefer_max_smem_carveout(K kern) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------- Philox4x32-10
// Counter = (env, site | sub-index, step_lo, step_hi); key = seed. One call yields 4 x 32 random bits.
enum DrawSite : uint32_t { kSiteQposNoise = 1, kSiteVelNoise = 2, kSiteResetF = 3, kSiteResetI = 4, kSitePert = 5, kSiteDR = 6 };

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}
__device__ __forceinline__ uint4 draw4(uint64_t seed, uint64_t step, uint32_t env, uint32_t site, uint32_t sub) {
  return philox4x32_10(make_uint4(env, (site << 24) | sub, (uint32_t)step, (uint32_t)(step >> 32)),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
// [0,1) with 24 random bits, the granularity of torch.rand(float32)
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// standard normal (Box-Muller) from two words
__device__ __forceinline__ float normal01(uint32_t a, uint32_t b) {
  float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0,1]
  float u2 = u01(b);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// standard normal pair (Box-Muller, both outputs) from two words
__device__ __forceinline__ float2 normal01_pair(uint32_t a, uint32_t b) {
  float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0,1]
  float u2 = u01(b);
  float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

// Out-of-line draw + Box-Muller for call sites that are unrolled several times (keeps the instruction footprint small).
static __device__ __noinline__ float2 draw_normal_pair(uint64_t seed, uint64_t step, uint32_t env, uint32_t site, uint32_t sub) {
  uint4 r = draw4(seed, step, env, site, sub);
  return normal01_pair(r.x, r.y);
}

// Division of small non-negative ints by a run-time constant: q = (i * mul) >> 20, exact while i * d < 2^20.
struct FastDiv {
  unsigned mul;
  int d;
  __host__ __device__ explicit FastDiv(int div) : mul(((1u << 20) + (unsigned)div - 1) / (unsigned)div), d(div) {}
  __device__ __forceinline__ int div(int i) const { return (int)(((unsigned)i * mul) >> 20); }
};

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// torch.remainder semantics for float (result takes the sign of the divisor)
__device__ __forceinline__ float py_fmodf(float a, float b) {
  float r = fmodf(a, b);
  if (r != 0.0f && ((r < 0.0f) != (b < 0.0f))) r += b;
  return r;
}

}  // namespace dyros
