// Stages of the per-step task logic that run both as their own kernels (task_kernels.cu, one warp per env) and
// inside the fused physics kernel (physics_kernels.cu, DYROS_LANES lanes per env). Arithmetic uses the
// non-contracting intrinsics so that both compilation units (with and without -fmad) produce the same bits,
// and the same float32 results as the reference's op-by-op torch code.
#pragma once
#include "internal.h"

namespace dyros {

struct TK {
  TaskParams p;
  DyrosTaskBuffers b;
  DyrosSimBuffers s;
  DyrosNoiseInjection j;
};

// T:505-520: upper-body PD, actuation-delay ring, the 33 torques of set_dof_actuation_force_tensor.
// LANES lanes of one env cooperate; `lane` in [0, LANES). Ends with the lanes in sync.
template <int LANES, class Sync>
__device__ __forceinline__ void stage_substep_torque(const TK& k, int e, int lane, Sync& sync) {
  const float* ds = k.s.dof_state + (size_t)e * ND * 2;
  float* out = k.s.dof_actuation_force + (size_t)e * ND;
  for (int d = 12 + lane; d < ND; d += LANES) {                                   // T:506
    float pos = ds[2 * d], vel = ds[2 * d + 1];
    out[d] = __fadd_rn(__fmul_rn(k.p.kp[d], __fsub_rn(k.b.target_data_qpos[(size_t)e * ND + d], pos)),
                       __fmul_rn(k.p.kv[d], -vel));
  }
  int sl = k.b.simul_len[e] + 1;                                                  // T:513-514
  sl = sl > LOG_DEPTH ? LOG_DEPTH : (sl < 0 ? 0 : sl);
  int dl = k.b.delay_idx[e];
  for (int j = lane; j < 12; j += LANES) {
    float* lg = k.b.action_log + (size_t)e * LOG_DEPTH * 12 + j;
    float v[LOG_DEPTH];
#pragma unroll
    for (int i = 0; i < LOG_DEPTH - 1; ++i) v[i] = lg[(i + 1) * 12];              // T:511
    v[LOG_DEPTH - 1] = k.b.action_torque[(size_t)e * 12 + j];                     // T:512
#pragma unroll
    for (int i = 0; i < LOG_DEPTH; ++i) lg[i * 12] = v[i];
    int pick = (sl > dl) ? dl : (LOG_DEPTH - sl);                                 // T:515-519
    float r = v[0];
#pragma unroll
    for (int i = 1; i < LOG_DEPTH; ++i) r = (pick == i) ? v[i] : r;
    out[j] = r;                                                                   // T:520
  }
  sync();
  if (lane == 0) k.b.simul_len[e] = sl;
}

// T:528-530
template <int LANES>
__device__ __forceinline__ void stage_sensor_noise(const TK& k, int substep, int e, int lane) {
  const float* ds = k.s.dof_state + (size_t)e * ND * 2;
  for (int d = lane; d < ND; d += LANES) {
    float n;
    if (k.j.qpos_normal) {
      n = k.j.qpos_normal[((size_t)substep * k.p.N + e) * ND + d];
    } else {
      uint4 r = draw4(k.p.seed, *k.p.step_counter, e, kSiteQposNoise, substep * 64 + d);
      n = __fmul_rn(normal01(r.x, r.y), k.p.noise_std);
    }
    n = (n < -0.00016f) ? -0.00016f : ((n > 0.00016f) ? 0.00016f : n);
    float qn = __fadd_rn(ds[2 * d], n);
    size_t i = (size_t)e * ND + d;
    k.b.qvel_noise[i] = __fdiv_rn(__fsub_rn(qn, k.b.qpos_pre[i]), k.p.dt);
    k.b.qpos_noise[i] = qn;
    k.b.qpos_pre[i] = qn;
  }
}

}  // namespace dyros
