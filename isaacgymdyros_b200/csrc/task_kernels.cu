// K2/K3/K4: the bodies of DyrosDynamicWalk's per-step methods as sm_100a kernels (one warp per env).
//
// Compiled with -fmad=false: the reference computes every float32 op separately (CPU torch / op-by-op
// ATen), and parity is judged at 1e-5 relative with bit-exact masks, so no FMA contraction here.
// Citations: T = tasks/dyros_dynamic_walk.py, VT = tasks/base/vec_task.py, JU = utils/torch_jit_utils.py,
// TU = isaacgym/torch_utils.py (all under the reference's python/ tree).
#include "task_stages.cuh"

namespace dyros {

constexpr int kWarpsPerBlock = 8;

// ------------------------------------------------------------------ small math (TU / JU restated)
// JU:142-160 with a = identity: |vec(a (x) conj(q))| through the TU:20-40 product, then 2*asin(min(.,1))
__device__ __forceinline__ float quat_err_identity(const float* q) {
  float x2 = -q[0], y2 = -q[1], z2 = -q[2], w2 = q[3];
  const float x1 = 0.f, y1 = 0.f, z1 = 0.f, w1 = 1.f;
  float ww = (z1 + x1) * (x2 + y2);
  float yy = (w1 - y1) * (w2 + z2);
  float zz = (w1 + y1) * (w2 - z2);
  float xx = ww + yy + zz;
  float qq = 0.5f * (xx + (z1 - x1) * (x2 - y2));
  float x = qq - xx + (x1 + w1) * (x2 + w2);
  float y = qq - yy + (w1 - x1) * (y2 + z2);
  float z = qq - zz + (z1 + y1) * (w2 - x2);
  float n = sqrtf(x * x + y * y + z * z);
  return 2.0f * asinf(fminf(n, 1.0f));
}

// TU:228-273 quat2euler (xyzw) -> component `which` (0:x 1:y 2:z)
__device__ __forceinline__ float quat2euler_comp(const float* q, int which) {
  float w = q[3], x = q[0], y = q[1], z = q[2];
  float m00 = w * w + x * x - y * y - z * z;
  float m01 = 2.f * x * y - 2.f * w * z;
  float m10 = 2.f * x * y + 2.f * w * z;
  float m11 = w * w - x * x + y * y - z * z;
  float m20 = 2.f * x * z - 2.f * w * y;
  float m21 = 2.f * y * z + 2.f * w * x;
  float m22 = w * w - x * x - y * y + z * z;
  float cy = sqrtf(m00 * m00 + m10 * m10);
  bool cond = cy > (float)(2.220446049250313e-16 * 4);
  // one atan2f for whichever component the lane wants (lanes with different `which` do not diverge):
  // 2: cond ? atan2(m10, m00) : atan2(-m01, m11); 1: atan2(-m20, cy); 0: cond ? atan2(m21, m22) : 0 (= atan2(0, 1))
  float ay = which == 2 ? (cond ? m10 : -m01) : (which == 1 ? -m20 : (cond ? m21 : 0.0f));
  float ax = which == 2 ? (cond ? m00 : m11) : (which == 1 ? cy : (cond ? m22 : 1.0f));
  return atan2f(ay, ax);
}

__device__ __forceinline__ bool collision_true(const TK& k, int e, int lane) {
  // T:590 / T:937: any non-foot body with |F| > 1
  bool hit = false;
  for (int b = lane; b < NB; b += kWarp) {
    if (b == k.p.lfoot || b == k.p.rfoot) continue;
    const float* f = k.s.net_contact_force + ((size_t)e * NB + b) * 3;
    float n = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    hit |= n > 1.0f;
  }
  return __any_sync(kFull, hit);
}

// ------------------------------------------------------------------ stages (device functions, one warp per env)
struct WarpSyncT {
  __device__ __forceinline__ void operator()() const { __syncwarp(); }
};

// T:532-541, VT:325, T:544-545
__device__ void stage_epilogue(const TK& k, int e, int lane) {
  if (lane != 0) return;
  k.b.epi_len[e] = k.b.epi_len[e] + 1.0f;
  float t = k.b.time[e] + k.p.dt_policy;
  t = t + k.p.time_gain * k.b.actions[(size_t)e * NA + NA - 1];
  k.b.time[e] = t;
  long long pr = k.b.progress_buf[e];
  k.b.timeout_buf[e] = ((float)pr >= k.p.max_len_m1) ? 1 : 0;
  k.b.progress_buf[e] = pr + 1;
  k.b.randomize_buf[e] = k.b.randomize_buf[e] + 1;
}

// T:581-596 ; returns the reset flag (warp-uniform)
// `shared` (fused kernel): hands the collision flag and the orientation error on to the reward stage, which needs
// the same two quantities of the same (pre-reset) state.
struct TermShared {
  bool col;
  float qe;
};
__device__ int stage_check_termination(const TK& k, int e, int lane, TermShared* shared = nullptr) {
  float qe = quat_err_identity(k.s.root_states + (size_t)e * 13 + 3);
  int reset = (fabsf(qe) > 0.5f) ? 1 : 0;
  if ((float)k.b.progress_buf[e] >= k.p.max_len_m1) reset = 1;
  const bool col = collision_true(k, e, lane);
  if (shared) {
    shared->col = col;
    shared->qe = qe;
  }
  if (col) reset = 1;
  if (lane == 0) k.b.reset_buf[e] = reset;
  return reset;
}

// T:387-428 + T:802-947, in two parts: the warp-wide reductions over the joints, and the scalar terms (one thread per
// env: the fused kernel runs them for all envs of a CTA side by side in one warp instead of on lane 0 of each).
struct RewardSums {
  float s_qpos, s_qvel, s_qacc, s_tq, s_tqd, qe;
  bool col;
};
__device__ RewardSums reward_sums(const TK& k, int e, int lane, const TermShared* shared = nullptr) {
  const float* root = k.s.root_states + (size_t)e * 13;
  const float* ds = k.s.dof_state + (size_t)e * ND * 2;
  float s_qpos = 0.f, s_qvel = 0.f, s_qacc = 0.f, s_tq = 0.f, s_tqd = 0.f;
  for (int d = lane; d < ND; d += kWarp) {
    float pos = ds[2 * d], vel = ds[2 * d + 1];
    float a = k.b.target_data_qpos[(size_t)e * ND + d] - pos;
    float b = 0.0f - vel;
    float c = vel - k.b.pre_joint_velocity_states[(size_t)e * ND + d];
    s_qpos += a * a;
    s_qvel += b * b;
    s_qacc += c * c;
  }
  if (lane < 12) {
    float a = k.b.actions[(size_t)e * NA + lane], ap = k.b.actions_pre[(size_t)e * NA + lane];
    float t = a * 333.0f, td = (a - ap) * 333.0f;
    s_tq = t * t;
    s_tqd = td * td;
  }
  RewardSums o;
  o.s_qpos = warp_sum(s_qpos);
  o.s_qvel = warp_sum(s_qvel);
  o.s_qacc = warp_sum(s_qacc);
  o.s_tq = warp_sum(s_tq);
  o.s_tqd = warp_sum(s_tqd);
  o.col = shared ? shared->col : collision_true(k, e, lane);
  o.qe = shared ? shared->qe : quat_err_identity(root + 3);
  return o;
}
__device__ void reward_scalar(const TK& k, int e, const RewardSums& in) {
  const float* root = k.s.root_states + (size_t)e * 13;
  const float s_qpos = in.s_qpos, s_qvel = in.s_qvel, s_qacc = in.s_qacc, s_tq = in.s_tq, s_tqd = in.s_tqd, qe = in.qe;
  const bool col = in.col;
  float r[14];
  r[0] = 0.3f * expf(-13.2f * fabsf(qe));                                          // T:835
  float n;
  n = sqrtf(s_qpos); r[1] = 0.35f * expf(-2.0f * (n * n));                         // T:837
  n = sqrtf(s_qvel); r[2] = 0.05f * expf(-0.01f * (n * n));                        // T:839
  const float* F = k.s.net_contact_force + (size_t)e * NB * 3;
  const float* Fp = k.b.contact_forces_pre + (size_t)e * NB * 3;
  const float *lf = F + k.p.lfoot * 3, *rf = F + k.p.rfoot * 3, *lfp = Fp + k.p.lfoot * 3, *rfp = Fp + k.p.rfoot * 3;
  float dl0 = lf[0] - lfp[0], dl1 = lf[1] - lfp[1], dl2 = lf[2] - lfp[2];
  float dr0 = rf[0] - rfp[0], dr1 = rf[1] - rfp[1], dr2 = rf[2] - rfp[2];
  r[9] = 0.2f * expf(-0.01f * (sqrtf(dl0 * dl0 + dl1 * dl1 + dl2 * dl2) + sqrtf(dr0 * dr0 + dr1 * dr1 + dr2 * dr2)));  // T:858
  r[4] = 0.05f * expf(-0.01f * sqrtf(s_tq));                                        // T:861
  r[5] = 0.6f * expf(-0.01f * sqrtf(s_tqd));                                        // T:863
  n = sqrtf(s_qacc); r[7] = 0.05f * expf(-20.0f * (n * n));                        // T:865
  float vx = k.b.target_vel[(size_t)e * 2] - root[7], vy = k.b.target_vel[(size_t)e * 2 + 1] - root[8];
  n = sqrtf(vx * vx + vy * vy); r[6] = 0.3f * expf(-3.0f * (n * n));               // T:867
  bool lc = lf[2] > 1.0f, rc = rf[2] > 1.0f;                                       // T:869-870
  int idx = k.b.mocap_data_idx[e];
  bool DSP = ((3300 <= idx) && (idx < 3600)) || (idx < 300) || ((1500 <= idx) && (idx < 2100));
  bool RSSP = (300 <= idx) && (idx < 1500);
  bool LSSP = (2100 <= idx) && (idx < 3300);
  bool sync = (DSP && rc && lc) || (RSSP && rc && !lc) || (LSSP && !rc && lc);     // T:882-889
  r[8] = sync ? 0.2f : 0.0f;
  k.b.contact_reward_sum[e] = k.b.contact_reward_sum[e] + r[8];                    // T:891
  r[10] = 0.0f;                                                                    // T:893
  float m = k.b.total_mass[e];
  float thr = (float)(1.4 * 9.81) * m;                                             // T:895
  bool thres = (lf[2] > thr) || (rf[2] > thr);
  r[11] = thres ? -0.2f : 0.0f;                                                    // T:898
  float cl = fmaxf(lf[2] - thr, 0.0f), cr = fmaxf(rf[2] - thr, 0.0f);
  float pen = 0.1f * expf(-0.007f * (sqrtf(cl * cl) + sqrtf(cr * cr)));            // T:900-901
  r[3] = thres ? pen : 0.1f;                                                       // T:902
  float dthr = (float)(0.2 * 9.81) * m / 1.0f;                                     // T:904
  bool tdiff = (fabsf(lf[2] - lfp[2]) > dthr) || (fabsf(rf[2] - rfp[2]) > dthr);
  r[12] = tdiff ? -0.05f : 0.0f;                                                   // T:907
  float ws = m / 104.48f;                                                          // T:917
  const float* ft = k.b.target_data_force + (size_t)e * 2;
  r[13] = 0.1f * expf(-0.001f * fabsf(lf[2] + ws * ft[0])) + 0.1f * expf(-0.001f * fabsf(rf[2] + ws * ft[1]));  // T:918-919
  float total = r[0] + r[1] + r[2] + r[3] + r[4] + r[5] + r[6] + r[7] + r[8] + r[9] + r[10] + r[11] + r[12] + r[13];  // T:932-934
  if (col) total = k.p.death_cost;                                                 // T:942
  if (fabsf(qe) > 0.5f) total = k.p.death_cost;                                    // T:943
  k.b.rew_buf[e] = total;
  float* st = k.b.stacked_rewards + (size_t)e * 15;
#pragma unroll
  for (int i = 0; i < 14; ++i) st[i] = col ? k.p.death_cost : r[i];                // T:945
  st[14] = (*k.b.perturb_start) ? 1.0f : 0.0f;                                     // T:415
}
__device__ void stage_compute_reward(const TK& k, int e, int lane, const TermShared* shared = nullptr) {
  RewardSums sums = reward_sums(k, e, lane, shared);
  if (lane == 0) reward_scalar(k, e, sums);
}

// T:598-669, T:720-748 (+ DR re-draw of damping/armature, VT:519-733 / gymutil.py:584-619)
__device__ void stage_reset_env(const TK& k, int e, int lane) {
  const TaskParams& P = k.p;
  // Philox epoch of this reset: the step epoch, plus the env's own count of explicit resets since the last step so that
  // an env reset twice within one epoch (reset_done() then a termination; repeated reset_idx calls) draws afresh
  const uint64_t step = *P.step_counter + ((uint64_t)k.b.reset_seq[e] << 40);
  // VT:540-544: only envs whose randomize_buf reached the frequency (1) are re-randomised; the buffer is cleared below
  const bool dr = P.randomize && k.b.randomize_buf[e] >= 1;
  __syncwarp();
  // --- draws (env-indexed). reset_f columns: see DyrosNoiseInjection
  auto uf = [&](int col) -> float {
    if (k.j.reset_f) return k.j.reset_f[(size_t)e * 32 + col];
    uint4 r = draw4(P.seed, step, e, kSiteResetF, col >> 2);
    uint32_t w = (col & 3) == 0 ? r.x : (col & 3) == 1 ? r.y : (col & 3) == 2 ? r.z : r.w;
    return u01(w);
  };
  for (int d = lane; d < ND; d += kWarp) {
    size_t i = (size_t)e * ND + d;
    k.b.qpos_noise[i] = P.init_dof_pos[d];                                         // T:611
    k.b.qpos_pre[i] = P.init_dof_pos[d];                                           // T:612
    k.b.qvel_noise[i] = 0.f;                                                       // T:613
    k.s.dof_state[2 * i] = P.reset_dof_pos[d];                                     // T:742
    k.s.dof_state[2 * i + 1] = 0.f;                                                // T:743
    k.b.pre_joint_velocity_states[i] = 0.f;                                        // T:635
    if (dr) {                                                                      // CFG:103-115, re-drawn on reset
      float u0, u1;
      if (k.j.dr_u) {
        u0 = k.j.dr_u[(size_t)e * 66 + d];
        u1 = k.j.dr_u[(size_t)e * 66 + 33 + d];
      } else {
        uint4 r = draw4(P.seed, step, e, kSiteDR, d);
        u0 = u01(r.x);
        u1 = u01(r.y);
      }
      k.s.dof_damping[i] = P.dr_damping_base + (P.dr_damping_lo + u0 * (P.dr_damping_hi - P.dr_damping_lo));
      k.s.dof_armature[i] = P.armature_base[d] * (P.dr_armature_lo + u1 * (P.dr_armature_hi - P.dr_armature_lo));
    }
  }
  // optional tables of the same randomisation pass (DyrosDynamicWalk.yaml:89-96 friction; PD gains: BASELINE configs[3]);
  // with injected DR draws (tests) they are left alone
  if (dr && !k.j.dr_u && lane == 0) {
    uint4 r = draw4(P.seed, step, e, kSiteDR, 64);
    if (k.s.contact_friction && P.dr_friction_hi > 0.f)
      k.s.contact_friction[e] = P.dr_friction_base * (P.dr_friction_lo + u01(r.x) * (P.dr_friction_hi - P.dr_friction_lo));
    if (k.b.pd_gain_scale && P.dr_pd_gain_hi > 0.f) {
      k.b.pd_gain_scale[(size_t)e * 2] = P.dr_pd_gain_lo + u01(r.y) * (P.dr_pd_gain_hi - P.dr_pd_gain_lo);
      k.b.pd_gain_scale[(size_t)e * 2 + 1] = P.dr_pd_gain_lo + u01(r.z) * (P.dr_pd_gain_hi - P.dr_pd_gain_lo);
    }
  }
  if (lane < 12) {
    k.b.qpos_bias[(size_t)e * 12 + lane] = uf(lane) * 6.28f / 100.f - (float)(3.14 / 100);        // T:615
    k.b.motor_constant_scale[(size_t)e * 12 + lane] = uf(18 + lane) * 0.4f + 0.8f;                // T:645
    k.b.action_torque_pre[(size_t)e * 12 + lane] = 0.f;                                            // T:636
  }
  if (lane < 3) k.b.quat_bias[(size_t)e * 3 + lane] = uf(12 + lane) * 6.28f / 150.f - (float)(3.14 / 150);  // T:616
  if (lane < 13) {                                                                 // T:734-735
    float v = 0.f;
    if (lane < 3) v = k.b.env_origins[(size_t)e * 3 + lane];
    if (lane == 2) v = P.initial_height + v;
    if (lane == 6) v = 1.f;
    k.s.root_states[(size_t)e * 13 + lane] = v;
  }
  for (int i = lane; i < NB * 3; i += kWarp)                                       // T:638
    k.b.contact_forces_pre[(size_t)e * NB * 3 + i] = k.s.net_contact_force[(size_t)e * NB * 3 + i];
  for (int i = lane; i < LOG_DEPTH * 12; i += kWarp) k.b.action_log[(size_t)e * LOG_DEPTH * 12 + i] = 0.f;  // T:651
  for (int i = lane; i < NSLOT * NOBS1; i += kWarp) k.b.obs_history[(size_t)e * NSLOT * NOBS1 + i] = 0.f;   // T:668
  for (int i = lane; i < NSLOT * NA; i += kWarp) k.b.action_history[(size_t)e * NSLOT * NA + i] = 0.f;       // T:669
  if (lane == 0) {
    float vel_mag = uf(15) * 0.8f;                                                 // T:624
    float vel_theta = uf(16) * 0.0f;                                               // T:625
    k.b.target_vel[(size_t)e * 2] = vel_mag * cosf(vel_theta);                     // T:626-628
    k.b.target_vel[(size_t)e * 2 + 1] = vel_mag * sinf(vel_theta);
    k.b.init_mocap_data_idx[e] = (uf(17) > 0.5f) ? 0 : 1800;                       // T:630-632
    k.b.time[e] = 0.f;                                                             // T:642
    k.b.progress_buf[e] = 0;                                                       // T:648
    k.b.reset_buf[e] = 1;                                                          // T:649
    int delay, timing;
    if (k.j.reset_i) {
      delay = (int)k.j.reset_i[(size_t)e * 2];
      timing = (int)k.j.reset_i[(size_t)e * 2 + 1];
    } else {
      uint4 r = draw4(P.seed, step, e, kSiteResetI, 0);
      delay = P.delay_lo + (int)(r.x % (uint32_t)(P.delay_hi - P.delay_lo));       // T:652
      timing = (int)(r.y % (uint32_t)P.timing_hi);                                 // T:665
    }
    k.b.delay_idx[e] = delay;
    float el = k.b.epi_len[e];
    k.b.contact_reward_mean[e] = k.b.contact_reward_sum[e] / el;                   // T:654
    k.b.contact_reward_sum[e] = 0.f;                                               // T:655
    k.b.simul_len[e] = 0;                                                          // T:657
    k.b.epi_len_log[e] = el;                                                       // T:658
    k.b.epi_len[e] = 0.f;                                                          // T:659
    k.b.perturbation_count[e] = 0;                                                 // T:663
    k.b.pert_on[e] = 0;                                                            // T:664
    k.b.perturb_timing[e] = timing;                                                // T:665
    if (dr) k.b.randomize_buf[e] = 0;                                              // VT:540-544 (frequency 1)
    k.b.reset_seq[e] = k.b.reset_seq[e] + 1;
  }
}

// T:750-796 (history kept as rings, see DyrosTaskBuffers)
__device__ void stage_compute_observations(const TK& k, int e, int lane) {
  const TaskParams& P = k.p;
  const float* root = k.s.root_states + (size_t)e * 13;
  float time = k.b.time[e];
  float time2idx = py_fmodf(time, P.period) / P.cycle_dt;                                           // T:762
  float phase = py_fmodf((float)k.b.init_mocap_data_idx[e] + time2idx, (float)P.mocap_data_num) / (float)P.mocap_data_num;  // T:763
  float ang = (float)(2 * 3.14159265358979) * phase;
  int head = (k.b.obs_hist_head[e] + 1) % NSLOT;
  bool start = k.b.epi_len[e] == 0.0f;                                                              // T:785
  float* hist = k.b.obs_history + (size_t)e * NSLOT * NOBS1;
  for (int i = lane; i < NOBS1; i += kWarp) {
    float v;
    if (i < 3) v = quat2euler_comp(root + 3, i) + k.b.quat_bias[(size_t)e * 3 + i];                 // T:753-757
    else if (i < 15) v = k.b.qpos_noise[(size_t)e * ND + (i - 3)] + k.b.qpos_bias[(size_t)e * 12 + (i - 3)];
    else if (i < 27) v = k.b.qvel_noise[(size_t)e * ND + (i - 15)];
    else if (i == 27) v = sinf(ang);                                                                // T:764
    else if (i == 28) v = cosf(ang);                                                                // T:765
    else if (i < 31) v = k.b.target_vel[(size_t)e * 2 + (i - 29)];
    else {
      int c = i - 31;
      float u;
      if (k.j.vel_u) u = k.j.vel_u[(size_t)e * 6 + c];
      else {
        uint4 r = draw4(P.seed, *P.step_counter, e, kSiteVelNoise, c);
        u = u01(r.x);
      }
      v = root[7 + c] + (u * 0.05f - 0.025f);                                                       // T:766,774
    }
    float nv = (v - k.b.obs_mean[i]) / sqrtf(k.b.obs_var[i] + 1e-8f * 1.0f);                        // T:776-777
    if (start) {
      for (int sl = 0; sl < NSLOT; ++sl) hist[sl * NOBS1 + i] = nv;                                 // T:786-787
    } else {
      hist[head * NOBS1 + i] = nv;                                                                  // T:783
    }
  }
  __syncwarp();
  if (lane == 0) k.b.obs_hist_head[e] = head;
  float* ob = k.b.obs_buf + (size_t)e * NOBS;
  const float* ah = k.b.action_history + (size_t)e * NSLOT * NA;
  int ahead = k.b.act_hist_head[e];
  // gather: all loads first (the compiler cannot reorder them across the stores itself), then the stores
  constexpr int IT = (NOBS + kWarp - 1) / kWarp;
  float vals[IT];
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    int o = lane + it * kWarp;
    o = o < NOBS ? o : NOBS - 1;
    float v;
    if (o < NOBS1 * NHIS) {                                                                         // T:789-791
      int i = o / NOBS1, j = o - i * NOBS1;
      int pos = NSKIP * (i + 1) - 1;
      v = hist[((head + 1 + pos) % NSLOT) * NOBS1 + j];
    } else {                                                                                        // T:793-796
      int q = o - NOBS1 * NHIS;
      int i = q / NA, j = q - i * NA;
      int pos = NSKIP * (i + 1);
      v = ah[((ahead + 1 + pos) % NSLOT) * NA + j];
    }
    vals[it] = v;
  }
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    int o = lane + it * kWarp;
    if (o < NOBS) ob[o] = vals[it];
  }
}

// T:560-563 (loads of all four copies are issued before the first store)
__device__ void stage_late_update(const TK& k, int e, int lane) {
  const float* ds = k.s.dof_state + (size_t)e * ND * 2;
  constexpr int ITC = (NB * 3 + kWarp - 1) / kWarp;
  float vel[2], cf[ITC], at = 0.f, ac = 0.f;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    int d = lane + it * kWarp;
    vel[it] = ds[2 * (d < ND ? d : ND - 1) + 1];
  }
#pragma unroll
  for (int it = 0; it < ITC; ++it) {
    int i = lane + it * kWarp;
    cf[it] = k.s.net_contact_force[(size_t)e * NB * 3 + (i < NB * 3 ? i : NB * 3 - 1)];
  }
  if (lane < 12) at = k.b.action_torque[(size_t)e * 12 + lane];
  if (lane < NA) ac = k.b.actions[(size_t)e * NA + lane];
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    int d = lane + it * kWarp;
    if (d < ND) k.b.pre_joint_velocity_states[(size_t)e * ND + d] = vel[it];
  }
  if (lane < 12) k.b.action_torque_pre[(size_t)e * 12 + lane] = at;
  if (lane < NA) k.b.actions_pre[(size_t)e * NA + lane] = ac;
#pragma unroll
  for (int it = 0; it < ITC; ++it) {
    int i = lane + it * kWarp;
    if (i < NB * 3) k.b.contact_forces_pre[(size_t)e * NB * 3 + i] = cf[it];
  }
}

// ------------------------------------------------------------------ kernels
#define ENV_LANE()                                                   \
  int lane = threadIdx.x & 31;                                       \
  int e = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);          \
  if (e >= k.p.N) return;

__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_prologue(TK k, const float* __restrict__ actions) {
  pdl_launch_dependents();
  ENV_LANE();
  stage_prologue(k, actions, e, lane);
}
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_substep_torque(TK k) {
  ENV_LANE();
  WarpSyncT sync;
  stage_substep_torque<32>(k, e, lane, sync);
}
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_sensor_noise(TK k, int substep) {
  ENV_LANE();
  stage_sensor_noise<32>(k, substep, e, lane);
}
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_epilogue(TK k) {
  ENV_LANE();
  stage_epilogue(k, e, lane);
}
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_check_termination(TK k) {
  ENV_LANE();
  stage_check_termination(k, e, lane);
}
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_compute_reward(TK k) {
  ENV_LANE();
  stage_compute_reward(k, e, lane);
}
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_reset_idx(TK k, const int64_t* __restrict__ ids, int count) {
  int lane = threadIdx.x & 31;
  int w = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  int n = count >= 0 ? count : *k.b.reset_count;
  if (w >= n || w >= k.p.N) return;
  long long e = ids[w];
  if (e < 0 || e >= k.p.N) return;
  stage_reset_env(k, (int)e, lane);
}
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_compute_observations(TK k) {
  ENV_LANE();
  stage_compute_observations(k, e, lane);
}
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_late_update(TK k) {
  ENV_LANE();
  stage_late_update(k, e, lane);
}
// one L2 prefetch per 128-byte line of [p, p + bytes), spread over the lanes of a warp
__device__ __forceinline__ void prefetch_rows_l2(const void* p, int bytes, int lane) {
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(p) & ~uintptr_t(127);
  const int lines = (int)((reinterpret_cast<uintptr_t>(p) + bytes + 127 - a0) >> 7);
  for (int i = lane; i < lines; i += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(a0 + (uintptr_t)i * 128));
}

// post_physics_step in one launch: epilogue, termination, reward, reset, observations, late update.
// Reward/termination read the pre-reset state, observations the post-reset state.
// The two means of the curriculum gate (T:489) are formed from fixed-point terms summed in 64-bit integers, so that the
// result does not depend on the order of summation: k_crossenv (one CTA, fixed order) and the fused post-physics launch
// (atomics across CTAs) give the same bits. epi_len_log holds whole numbers (exact at 2^-20), contact_reward_mean lies
// in [0, 1] (2^-40 is below a float's resolution there).
__device__ __forceinline__ long long gate_term0(float epi_len_log) { return __double2ll_rn((double)epi_len_log * 1048576.0); }
__device__ __forceinline__ long long gate_term1(float contact_reward_mean) { return __double2ll_rn((double)contact_reward_mean * 1099511627776.0); }
__device__ __forceinline__ bool gate_open(const TaskParams& P, long long s0, long long s1) {  // T:489
  return (float)((double)s0 / 1048576.0 / P.N) > P.gate_len && (float)((double)s1 / 1099511627776.0 / P.N) > 0.165f;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_post_fused(TK k, int tail) {
  __shared__ RewardSums sums[kWarpsPerBlock];
  __shared__ long long gate_acc[2][kWarpsPerBlock];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int e = blockIdx.x * kWarpsPerBlock + w;
  const bool valid = e < k.p.N;  // (no early return: the CTA meets at two barriers)
  int reset = 0;
  if (valid) {
    // The stages below are a chain of dependent reads. Once the env state no longer fits L2 (about 10.4 KB per env:
    // beyond ~8k envs per GPU) the rows the physics launch did not touch (history rings, previous-step copies) are
    // requested from HBM now, all at once, instead of one miss per stage (+6 % at 16,384 envs; at 4,096 envs they are
    // L2 hits and the extra instructions only cost).
    if (k.p.N > 8192) {
      prefetch_rows_l2(k.b.obs_history + (size_t)e * NSLOT * NOBS1, NSLOT * NOBS1 * 4, lane);
      prefetch_rows_l2(k.b.action_history + (size_t)e * NSLOT * NA, NSLOT * NA * 4, lane);
      prefetch_rows_l2(k.b.contact_forces_pre + (size_t)e * NB * 3, NB * 3 * 4, lane);
      prefetch_rows_l2(k.b.pre_joint_velocity_states + (size_t)e * ND, ND * 4, lane);
      prefetch_rows_l2(k.b.actions_pre + (size_t)e * NA, NA * 4, lane);
    }
    stage_epilogue(k, e, lane);
    __syncwarp();
    TermShared ts;
    reset = stage_check_termination(k, e, lane, &ts);
    __syncwarp();
    RewardSums rs = reward_sums(k, e, lane, &ts);
    if (lane == 0) sums[w] = rs;
  }
  __syncthreads();
  if (w == 0 && lane < kWarpsPerBlock) {  // the scalar reward terms of the CTA's envs, one lane per env
    const int e2 = blockIdx.x * kWarpsPerBlock + lane;
    if (e2 < k.p.N) reward_scalar(k, e2, sums[lane]);
  }
  __syncthreads();
  if (valid) {
    if (reset) {
      stage_reset_env(k, e, lane);
      __syncwarp();
    }
    stage_compute_observations(k, e, lane);
    __syncwarp();
    stage_late_update(k, e, lane);
  }
  if (!tail) return;
  // ---- cross-env pass of the fused step (what k_crossenv does for the staged one, minus the id list): gate sums by
  //      64-bit integer atomics, one pair per CTA; the last CTA to finish decides the gate and bumps the Philox epoch
  const bool gate = k.p.perturb != 0;
  __syncwarp();
  if (gate && lane == 0) {
    gate_acc[0][w] = valid ? gate_term0(k.b.epi_len_log[e]) : 0;          // (post-reset values, as k_crossenv reads them)
    gate_acc[1][w] = valid ? gate_term1(k.b.contact_reward_mean[e]) : 0;
  }
  // (no device-wide fence by every thread: nothing written above is read by another CTA of this launch; what the last
  //  CTA reads are the sums below, and what it overwrites, the Philox epoch, every warp has finished reading before
  //  the barrier; thread 0's fences order its atomics)
  __syncthreads();
  if (threadIdx.x == 0) {
    if (gate) {
      long long a = 0, b = 0;
#pragma unroll
      for (int i = 0; i < kWarpsPerBlock; ++i) {
        a += gate_acc[0][i];
        b += gate_acc[1][i];
      }
      atomicAdd(k.p.tail + 0, (unsigned long long)a);
      atomicAdd(k.p.tail + 1, (unsigned long long)b);
    }
    __threadfence();
    if (atomicAdd(k.p.tail + 2, 1ull) == gridDim.x - 1) {  // every CTA has finished its envs and published its sums
      __threadfence();
      if (gate) {
        const long long s0 = (long long)atomicExch(k.p.tail + 0, 0ull), s1 = (long long)atomicExch(k.p.tail + 1, 0ull);
        if (gate_open(k.p, s0, s1)) *k.b.perturb_start = 1;  // T:489-490 (sticky)
      }
      k.p.tail[2] = 0;
      *k.p.step_counter = *k.p.step_counter + 1;
    }
  }
}

// Cross-env pass (one block, one sweep over the envs): (a) reset_buf.nonzero() -> ascending ids (T:554): thread t owns the
// consecutive envs [t * per, (t + 1) * per), so a block-wide exclusive scan of the per-thread counts (warp shuffle scan,
// then a scan of the 32 warp totals) gives every thread the output position of its first id; (b) the curriculum gate
// means of T:489 (fixed summation order: deterministic); (c) Philox epoch bump.
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) k_crossenv(TK k, int do_compact, int do_gate, int do_bump) {
  __shared__ int warp_tot[kScanThreads / 32];
  __shared__ long long red[2][kScanThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int N = k.p.N;
  const int per = (N + kScanThreads - 1) / kScanThreads;
  const int e0 = tid * per, e1 = min(N, e0 + per);
  pdl_wait();
  int cnt = 0;
  long long s0 = 0, s1 = 0;  // fixed-point terms: see gate_term0 / gate_term1
  const bool gate = do_gate && k.p.perturb;
  for (int e = e0; e < e1; ++e) {
    if (do_compact) cnt += k.b.reset_buf[e] != 0;
    if (gate) {
      s0 += gate_term0(k.b.epi_len_log[e]);
      s1 += gate_term1(k.b.contact_reward_mean[e]);
    }
  }
  // inclusive scan of cnt inside the warp, warp totals to shared memory
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(kFull, s0, o);
    s1 += __shfl_xor_sync(kFull, s1, o);
  }
  if (lane == 31) warp_tot[w] = inc;
  if (lane == 0) {
    red[0][w] = s0;
    red[1][w] = s1;
  }
  __syncthreads();
  if (w == 0) {
    int v = warp_tot[lane], sc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(kFull, sc, o);
      if (lane >= o) sc += t;
    }
    warp_tot[lane] = sc - v;  // exclusive offsets of the warps
    if (lane == 31 && do_compact) *k.b.reset_count = sc;
    if (gate) {
      long long a = red[0][lane], b = red[1][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(kFull, a, o);
        b += __shfl_xor_sync(kFull, b, o);
      }
      if (lane == 0 && gate_open(k.p, a, b)) *k.b.perturb_start = 1;  // T:489-490 (sticky)
    }
    if (do_bump && lane == 0) *k.p.step_counter = *k.p.step_counter + 1;
  }
  __syncthreads();
  if (do_compact && cnt) {
    int pos = warp_tot[w] + inc - cnt;
    for (int e = e0; e < e1; ++e)
      if (k.b.reset_buf[e] != 0) {
        k.b.reset_env_ids[pos] = e;
        k.b.reset_env_ids32[pos] = e;
        ++pos;
      }
  }
}


// ------------------------------------------------------------------ results of a step -> one contiguous block
// Layout of `dst`: obs (N,487) f32 | rew (N) f32 | reset (N) i64 | time_outs (N) i64 (the i64 part starts at
// N*488*4, a multiple of 8). The host-facing
// step (DyrosDynamicWalk.step_async) hands this block to the copy engine as ONE device->host transfer on a second
// stream, so the next step may overwrite obs_buf / rew_buf / reset_buf / timeout_buf (VT:336-344) while the block is still in flight.
__global__ void __launch_bounds__(256) k_pack_results(TK k, float* __restrict__ dst) {
  const int N = k.p.N;
  const size_t n_obs = (size_t)N * 487;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  const size_t n4 = n_obs / 4;  // obs_buf and dst are 16-byte aligned (torch allocations)
  const float4* src4 = reinterpret_cast<const float4*>(k.b.obs_buf);
  float4* dst4 = reinterpret_cast<float4*>(dst);
  if (k.b.obs_buf != dst) {  // (dyros_task_set_obs_buf(dst): the observation kernel has written the block itself)
    // Both sides of this copy are touched once per step: L2 evict-first, so that the 16 MB streamed here do not linger
    // in L2 at the expense of the env state (measured: plain loads / stores cost the next step 75 us, with the hints
    // 50 us; the pipelined host step therefore avoids this branch altogether, profiles/r1j_step_async.txt)
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (size_t i = tid; i < n4; i += nth) {
      float4 v;
      asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src4 + i), "l"(pol));
      asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                   ::"l"(dst4 + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
    }
    for (size_t i = n4 * 4 + tid; i < n_obs; i += nth) dst[i] = k.b.obs_buf[i];
  }
  float* rew = dst + n_obs;
  long long* rst = reinterpret_cast<long long*>(dst + n_obs + N);
  for (size_t e = tid; e < (size_t)N; e += nth) {
    rew[e] = k.b.rew_buf[e];
    rst[e] = k.b.reset_buf[e];
    rst[N + e] = k.b.timeout_buf[e];
  }
}

// ------------------------------------------------------------------ launchers
static inline TK make_tk(Task* t) {
  TK k;
  k.p = t->p;
  k.b = t->b;
  k.s = t->sim->b;
  k.j = t->inj;
  return k;
}
static inline int env_grid(int N) { return (N + kWarpsPerBlock - 1) / kWarpsPerBlock; }
#define LAUNCH_ENV(kern, ...)                                                    \
  kern<<<env_grid(t->p.N), kWarpsPerBlock * 32, 0, s>>>(make_tk(t), ##__VA_ARGS__); \
  DY_LAUNCH_CHECK();                                                             \
  return 0;

int launch_prologue(Task* t, const float* actions, cudaStream_t s) { LAUNCH_ENV(k_prologue, actions) }
int launch_substep_torque(Task* t, cudaStream_t s) { LAUNCH_ENV(k_substep_torque) }
int launch_sensor_noise(Task* t, int substep, cudaStream_t s) { LAUNCH_ENV(k_sensor_noise, substep) }
int launch_epilogue(Task* t, cudaStream_t s) { LAUNCH_ENV(k_epilogue) }
int launch_check_termination(Task* t, cudaStream_t s) { LAUNCH_ENV(k_check_termination) }
int launch_compute_reward(Task* t, cudaStream_t s) { LAUNCH_ENV(k_compute_reward) }
int launch_compute_observations(Task* t, cudaStream_t s) { LAUNCH_ENV(k_compute_observations) }
int launch_late_update(Task* t, cudaStream_t s) { LAUNCH_ENV(k_late_update) }
int launch_post_fused(Task* t, cudaStream_t s, bool pdl, bool tail) {
  DY_CUDA(launch_kernel(k_post_fused, dim3(env_grid(t->p.N)), dim3(kWarpsPerBlock * 32), 0, s, pdl, make_tk(t), (int)tail));
  return 0;
}
int launch_reset_idx(Task* t, const int64_t* env_ids, int count, cudaStream_t s) {
  if (count == 0) return 0;
  const int64_t* ids = env_ids ? env_ids : t->b.reset_env_ids;
  int n = count >= 0 ? count : t->p.N;
  k_reset_idx<<<env_grid(n), kWarpsPerBlock * 32, 0, s>>>(make_tk(t), ids, count);
  DY_LAUNCH_CHECK();
  return 0;
}
int launch_pack_results(Task* t, float* dst, cudaStream_t s) {
  // a multiple of the SM count; 8 MB at N = 4096: ~13 float4 per thread
  k_pack_results<<<148 * 4, 256, 0, s>>>(make_tk(t), dst);
  DY_LAUNCH_CHECK();
  return 0;
}
int launch_crossenv(Task* t, bool compact, bool gate, bool bump, cudaStream_t s, bool pdl) {
  DY_CUDA(launch_kernel(k_crossenv, dim3(1), dim3(kScanThreads), 0, s, pdl, make_tk(t), (int)compact, (int)gate, (int)bump));
  return 0;
}

}  // namespace dyros
