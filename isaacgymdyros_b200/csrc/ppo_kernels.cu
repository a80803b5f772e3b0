// On-device PPO pieces around the env step: the bodies of the rollout bookkeeping, GAE, the clipped
// surrogate / value loss gradient and the two Adam optimisers of the reference's rl_games fork, as sm_100a kernels.
// The two 487-256-256-{13,1} MLPs stay library GEMMs (torch / cuBLAS); everything elementwise around them is here, so
// that one rollout step and one minibatch update are each a handful of launches inside a CUDA graph.
// Citations: A2C = learning/rl_games_custom/a2c_common_dyros.py, AG = learning/rl_games_custom/a2c_continuous_seperate.py,
// MD = learning/rl_games_custom/models_dyros.py, PPO = cfg/train/DyrosDynamicWalkPPO.yaml (reference tree). The loss
// formulas of rl_games 1.1.4 (common_losses.actor_loss / critic_loss, torch_ext.policy_kl; pinned by
// IsaacGymEnvs/setup.py:22, not vendored) are restated from their published definitions; oracle/ppo_oracle.py holds the
// plain-torch restatement the tests compare with.
#include <math.h>

#include "common.cuh"

namespace dyros {

constexpr int PPO_NA = 13, PPO_NOBS = 487;

// ---------------------------------------------------------------- rollout: act + record (A2C:629-647, MD:28-58)
// One warp per env. Samples a = mu + sigma * N(0,1) (Philox, counter = (env, global step)), its negative log
// likelihood (MD:60-63), and records obs / done / mu / value / action / neglogp at slot n = *step of the env-major
// rollout buffers, (N, H, .): a minibatch of the update is then a contiguous block of rows, as rl_games' dataset slices
// of the swap_and_flatten01 layout are (A2C:703, 32 envs x 128 steps for the reference's sizes).
__global__ void __launch_bounds__(256) k_ppo_act(DyrosPpoBuffers b, const float* __restrict__ mu, const float* __restrict__ value,
                                                 const float* __restrict__ logstd, const float* __restrict__ obs,
                                                 const long long* __restrict__ reset_buf, float* __restrict__ actions_env,
                                                 const float* __restrict__ inject_normal) {
  const int lane = threadIdx.x & 31, e = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (e >= b.N) return;
  const int n = *b.step;
  const size_t row = (size_t)e * b.H + n;
  float a = 0.f, z = 0.f, m = 0.f, ls = 0.f;
  if (lane < PPO_NA) {
    m = mu[(size_t)e * PPO_NA + lane];
    ls = logstd[lane];
    if (inject_normal) z = inject_normal[((size_t)n * b.N + e) * PPO_NA + lane];
    else {
      const uint4 r = draw4(b.seed, *b.global_step, e, 7u /* policy site */, lane >> 1);
      const float2 nn = normal01_pair(r.x, r.y);
      z = (lane & 1) ? nn.y : nn.x;
    }
    a = m + expf(ls) * z;  // torch.distributions.Normal(mu, sigma).sample(), MD:46-47
  }
  // neglogp = 0.5 * sum(((a - mu) / sigma)^2) + 0.5 * log(2 pi) * 13 + sum(logstd), MD:60-63
  float t = 0.f;
  if (lane < PPO_NA) {
    const float d = (a - m) / expf(ls);
    t = 0.5f * d * d + ls;
  }
  t = warp_sum(t) + 0.5f * 1.8378770664093453f * (float)PPO_NA;
  if (lane < PPO_NA) {
    b.actions[row * PPO_NA + lane] = a;
    b.mus[row * PPO_NA + lane] = m;
    actions_env[(size_t)e * PPO_NA + lane] = a;  // the env clamps to [-1, 1] itself (VT:307; A2C:820-824 rescale = identity)
  }
  if (lane == 0) {
    b.neglogp[row] = t;
    b.values[row] = value[e];
    b.dones[row] = reset_buf[e] != 0 ? 1.f : 0.f;  // self.dones BEFORE the step (A2C:640)
  }
  const float* src = obs + (size_t)e * PPO_NOBS;
  float* dst = b.obs + row * PPO_NOBS;
  for (int i = lane; i < PPO_NOBS; i += 32) dst[i] = src[i];
}

// after the env step: shaped reward with the time-out bootstrap (A2C:654-661), episode statistics (A2C:663-684)
__global__ void __launch_bounds__(256) k_ppo_reward(DyrosPpoBuffers b, const float* __restrict__ rew, const long long* __restrict__ timeout,
                                                    const long long* __restrict__ reset_buf) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = *b.step;
  if (e < b.N) {
    const size_t row = (size_t)e * b.H + n;
    float r = rew[e] * b.reward_scale;
    if (b.value_bootstrap && timeout[e] != 0) r += b.gamma * b.values[row];
    b.rewards[row] = r;
    const float cr = b.cur_reward[e] + rew[e], cl = b.cur_length[e] + 1.f;
    const bool done = reset_buf[e] != 0;
    if (done) {  // finished episodes feed the running means (game_rewards / game_lengths)
      atomicAdd(b.ep_stats + 0, cr);
      atomicAdd(b.ep_stats + 1, cl);
      atomicAdd(b.ep_stats + 2, 1.f);
    }
    b.cur_reward[e] = done ? 0.f : cr;
    b.cur_length[e] = done ? 0.f : cl;
  }
}
__global__ void k_ppo_advance(DyrosPpoBuffers b) {
  *b.step = (*b.step + 1) % b.H;
  *b.global_step = *b.global_step + 1;
}

// ---------------------------------------------------------------- GAE (A2C:485-500) + returns (A2C:692)
// One thread per env, walking its contiguous (H) rows backwards.
__global__ void __launch_bounds__(128) k_ppo_gae(DyrosPpoBuffers b, const float* __restrict__ last_values,
                                                 const long long* __restrict__ last_reset) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= b.N) return;
  const size_t r0 = (size_t)e * b.H;
  float lastgaelam = 0.f;
  float nextvalue = last_values[e];
  float nextnonterminal = 1.f - (last_reset[e] != 0 ? 1.f : 0.f);
  for (int t = b.H - 1; t >= 0; --t) {
    const float v = b.values[r0 + t];
    const float delta = b.rewards[r0 + t] + b.gamma * nextvalue * nextnonterminal - v;
    lastgaelam = delta + b.gamma * b.tau * nextnonterminal * lastgaelam;
    b.advantages[r0 + t] = lastgaelam;
    b.returns[r0 + t] = lastgaelam + v;
    nextvalue = v;
    nextnonterminal = 1.f - b.dones[r0 + t];
  }
}

// ---------------------------------------------------------------- loss gradient of one minibatch (AG:108-160)
// For rows [row0, row0 + mb): given the networks' outputs mu (mb,13) and value (mb), writes dLoss/dmu and dLoss/dvalue
// of  loss = mean(max(-A r, -A clip(r, 1-e, 1+e))) + 0.5 * critic_coef * mean((ret - v)^2)   (entropy and bound
// coefficients are 0 in PPO:80-92; sigma is a fixed, non-trainable parameter, network_builder_dyros.py:103) with
// r = exp(old_neglogp - neglogp(a | mu, sigma)), and accumulates the logged sums: actor loss, critic loss, KL
// (torch_ext.policy_kl), clip fraction. One warp per row.
__global__ void __launch_bounds__(256) k_ppo_loss_grad(DyrosPpoBuffers b, int row0, int mb, const float* __restrict__ mu,
                                                       const float* __restrict__ value, const float* __restrict__ logstd,
                                                       const float* __restrict__ adv_norm, float* __restrict__ dmu,
                                                       float* __restrict__ dvalue, float* __restrict__ stats) {
  __shared__ float acc[4][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * 8 + w;
  float s_a = 0.f, s_c = 0.f, s_kl = 0.f, s_cf = 0.f;
  if (i < mb) {
    const size_t row = (size_t)row0 + i;
    float m = 0.f, a = 0.f, ls = 0.f, om = 0.f, inv_s = 0.f;
    if (lane < PPO_NA) {
      m = mu[(size_t)i * PPO_NA + lane];
      a = b.actions[row * PPO_NA + lane];
      om = b.mus[row * PPO_NA + lane];
      ls = logstd[lane];
      inv_s = expf(-ls);
    }
    const float d = (a - m) * inv_s;
    float t = lane < PPO_NA ? 0.5f * d * d + ls : 0.f;
    const float neglogp = warp_sum(t) + 0.5f * 1.8378770664093453f * (float)PPO_NA;
    const float A = adv_norm[row];
    const float ratio = expf(b.neglogp[row] - neglogp);
    const float lo = 1.f - b.e_clip, hi = 1.f + b.e_clip;
    const float rc = fminf(fmaxf(ratio, lo), hi);
    const float l1 = -A * ratio, l2 = -A * rc;
    // d max(l1, l2) / d ratio: -A on the unclipped branch, -A * 1[lo < ratio < hi] on the clipped one
    const bool first = l1 >= l2;
    const float dl_dratio = first ? -A : ((ratio > lo && ratio < hi) ? -A : 0.f);
    // d ratio / d mu_k = ratio * (a_k - mu_k) / sigma_k^2
    const float inv_mb = 1.f / (float)mb;
    if (lane < PPO_NA) dmu[(size_t)i * PPO_NA + lane] = dl_dratio * ratio * d * inv_s * inv_mb;
    const float v = value[i], ret = b.returns[row];
    if (lane == 0) dvalue[i] = 0.5f * b.critic_coef * 2.f * (v - ret) * inv_mb;
    // KL(new || old) as torch_ext.policy_kl(p0 = new, p1 = old): sigma is shared, so c1 + c3 = log(1 + 1e-5) - 0.5
    float kl = 0.f;
    if (lane < PPO_NA) {
      const float s2 = expf(2.f * ls);
      kl = logf(1.f + 1e-5f) + (s2 + (om - m) * (om - m)) / (2.f * (s2 + 1e-5f)) - 0.5f;
    }
    s_kl = warp_sum(kl);
    s_a = fmaxf(l1, l2);
    s_c = (ret - v) * (ret - v);
    s_cf = fabsf(ratio - 1.f) > b.e_clip ? 1.f : 0.f;
  }
  if (lane == 0) {
    acc[0][w] = s_a; acc[1][w] = s_c; acc[2][w] = s_kl; acc[3][w] = s_cf;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += acc[threadIdx.x][k];
    atomicAdd(stats + threadIdx.x, s / (float)mb);
  }
}

// ---------------------------------------------------------------- the two Adam optimisers in one pass (AG:50-54, AG:164-187)
// Flat parameter / gradient / moment buffers: [0, n_actor) = actor (clip_grad_norm_(actor_param, grad_norm), AG:179),
// [n_actor, n) = critic (not clipped). grads are SUMS over ranks: grad_scale = 1 / world averages them (the reference's
// Horovod DistributedOptimizer averages, AG:161-163). lr and the step count live on the device (graph replays).
__global__ void __launch_bounds__(256) k_ppo_gradnorm(const float* __restrict__ g, int n_actor, float grad_scale, float* __restrict__ norm2) {
  float s = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_actor; i += gridDim.x * blockDim.x) {
    const float x = g[i] * grad_scale;
    s += x * x;
  }
  s = warp_sum(s);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += part[k];
    atomicAdd(norm2, t);
  }
}
__global__ void __launch_bounds__(256) k_ppo_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                  float* __restrict__ v, int n_actor, int n, float grad_scale, float max_norm,
                                                  float* __restrict__ norm2, const float* __restrict__ lr_dev, int* __restrict__ t_dev,
                                                  float beta1, float beta2, float eps) {
  const int t = *t_dev + 1;
  const float lr_actor = lr_dev[0], lr_critic = lr_dev[1];  // AG:53-54: the critic's rate is fixed, the actor's is scheduled
  const float norm = sqrtf(*norm2);
  const float clip = fminf(max_norm / (norm + 1e-6f), 1.f);  // torch.nn.utils.clip_grad_norm_
  const float bc1 = 1.f - powf(beta1, (float)t), bc2 = 1.f - powf(beta2, (float)t);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    if (i < n_actor && max_norm > 0.f) gi *= clip;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= (i < n_actor ? lr_actor : lr_critic) * (mi / bc1) / (sqrtf(vi / bc2) + eps);  // torch.optim.Adam, amsgrad off, weight_decay 0
  }
}
// ... and rl_games' LinearScheduler, which the reference steps once per MINIBATCH (schedule_type 'legacy', A2C:888-892;
// max_steps = max_epochs, A2C:135-137): lr = min_lr + (lr0 - min_lr) * max(0, max_steps - steps) / max_steps
__global__ void k_ppo_adam_finish(float* norm2, int* t_dev, float* lr_dev, float lr0, float lr_min, int max_steps) {
  *norm2 = 0.f;
  const int t = *t_dev + 1;
  *t_dev = t;
  if (max_steps > 0) {
    const int left = max_steps - t > 0 ? max_steps - t : 0;
    lr_dev[0] = lr_min + (lr0 - lr_min) * ((float)left / (float)max_steps);
  }
}

}  // namespace dyros

using namespace dyros;
void dyros_set_error_ppo(const char* msg) { dyros::set_error("%s", msg); }
#define PPO_CHECK(cond, msg)   \
  do {                         \
    if (!(cond)) {             \
      dyros_set_error_ppo(msg); \
      return 1;                \
    }                          \
  } while (0)

namespace dyros {
static int configure_ppo_kernels() {  // (common.cuh: one carve-out for every kernel of the library)
  static bool done = false;
  if (done) return 0;
  DY_CUDA(prefer_max_smem_carveout(k_ppo_act)); DY_CUDA(prefer_max_smem_carveout(k_ppo_reward));
  DY_CUDA(prefer_max_smem_carveout(k_ppo_advance)); DY_CUDA(prefer_max_smem_carveout(k_ppo_gae));
  DY_CUDA(prefer_max_smem_carveout(k_ppo_loss_grad)); DY_CUDA(prefer_max_smem_carveout(k_ppo_gradnorm));
  DY_CUDA(prefer_max_smem_carveout(k_ppo_adam)); DY_CUDA(prefer_max_smem_carveout(k_ppo_adam_finish));
  done = true;
  return 0;
}
}  // namespace dyros

extern "C" {

int dyros_ppo_act(const DyrosPpoBuffers* b, const float* mu, const float* value, const float* logstd, const float* obs,
                  const int64_t* reset_buf, float* actions_env, const float* inject_normal, void* stream) {
  PPO_CHECK(b && mu && value && logstd && obs && reset_buf && actions_env, "dyros_ppo_act: null argument");
  if (configure_ppo_kernels()) return 1;
  k_ppo_act<<<(b->N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(*b, mu, value, logstd, obs, reinterpret_cast<const long long*>(reset_buf),
                                                              actions_env, inject_normal);
  DY_LAUNCH_CHECK();
  return 0;
}
int dyros_ppo_reward(const DyrosPpoBuffers* b, const float* rew, const int64_t* timeout, const int64_t* reset_buf, void* stream) {
  PPO_CHECK(b && rew && timeout && reset_buf, "dyros_ppo_reward: null argument");
  if (configure_ppo_kernels()) return 1;
  k_ppo_reward<<<(b->N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*b, rew, reinterpret_cast<const long long*>(timeout),
                                                                    reinterpret_cast<const long long*>(reset_buf));
  k_ppo_advance<<<1, 1, 0, (cudaStream_t)stream>>>(*b);
  DY_LAUNCH_CHECK();
  return 0;
}
int dyros_ppo_gae(const DyrosPpoBuffers* b, const float* last_values, const int64_t* last_reset, void* stream) {
  PPO_CHECK(b && last_values && last_reset, "dyros_ppo_gae: null argument");
  if (configure_ppo_kernels()) return 1;
  k_ppo_gae<<<(b->N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*b, last_values, reinterpret_cast<const long long*>(last_reset));
  DY_LAUNCH_CHECK();
  return 0;
}
int dyros_ppo_loss_grad(const DyrosPpoBuffers* b, int row0, int mb, const float* mu, const float* value, const float* logstd,
                        const float* adv_norm, float* dmu, float* dvalue, float* stats, void* stream) {
  PPO_CHECK(b && mu && value && logstd && adv_norm && dmu && dvalue && stats, "dyros_ppo_loss_grad: null argument");
  if (configure_ppo_kernels()) return 1;
  PPO_CHECK(row0 >= 0 && mb > 0 && (long long)row0 + mb <= (long long)b->N * b->H, "dyros_ppo_loss_grad: rows outside the rollout");
  k_ppo_loss_grad<<<(mb + 7) / 8, 256, 0, (cudaStream_t)stream>>>(*b, row0, mb, mu, value, logstd, adv_norm, dmu, dvalue, stats);
  DY_LAUNCH_CHECK();
  return 0;
}
int dyros_ppo_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int n_actor, int n, float grad_scale,
                   float max_norm, float* norm2_scratch, float* lr_dev, int32_t* step_dev, float beta1, float beta2,
                   float eps, float lr0, float lr_min, int lr_max_steps, void* stream) {
  PPO_CHECK(params && grads && exp_avg && exp_avg_sq && norm2_scratch && lr_dev && step_dev, "dyros_ppo_adam: null argument");
  if (configure_ppo_kernels()) return 1;
  PPO_CHECK(n_actor >= 0 && n_actor <= n, "dyros_ppo_adam: n_actor outside [0, n]");
  cudaStream_t s = (cudaStream_t)stream;
  if (max_norm > 0.f && n_actor > 0) k_ppo_gradnorm<<<148, 256, 0, s>>>(grads, n_actor, grad_scale, norm2_scratch);
  k_ppo_adam<<<148 * 2, 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, n_actor, n, grad_scale, max_norm, norm2_scratch, lr_dev,
                                    step_dev, beta1, beta2, eps);
  k_ppo_adam_finish<<<1, 1, 0, s>>>(norm2_scratch, step_dev, lr_dev, lr0, lr_min, lr_max_steps);
  DY_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
