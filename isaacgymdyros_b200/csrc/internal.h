// Internal (non-ABI) structures shared by capi.cu, task_kernels.cu and physics_kernels.cu.
#pragma once
#include "common.cuh"

namespace dyros {

constexpr int ND = 33, NB = 38, NA = 13, NOBS1 = 37, NHIS = 10, NSKIP = 2, NSLOT = NHIS * NSKIP;  // T:36-43
constexpr int NOBS = (NOBS1 + NA) * (NHIS - 1) + NOBS1;                                           // 487
constexpr int LOG_DEPTH = 6;                                                                      // T:166

// float32 device copies of the model tables (see model/tables.py for the meaning of each array)
struct DevModel {
  int nl, nb, nd, np, nc, ns, T;
  const int* link_parent;
  const int* link_dof;
  const float* link_E;
  const float* link_r;
  const float* link_axis;
  const int* body_link;
  const float* body_pos;
  const float* body_rot;
  const float* body_inertia;
  const float* dof_lower;
  const float* dof_upper;
  const float* dof_vel_limit;
  const float* dof_effort;
  const int* pt_link;
  const int* pt_body;
  const float* pt_pos;
  const float* pt_radius;
  const int* cyl_link;
  const int* cyl_body;
  const float* cyl_center;
  const float* cyl_axis;
  const float* cyl_size;
  const int* solver_links;
  const int* sched;
  const int* link_solver_slot;  // [nl] index into solver_links or -1
};

struct SimParams {
  int N;
  float dt;
  int substeps;
  float g[3];
  float contact_offset, max_depen_vel, mu, pen_k, pen_c, max_ang_vel;
  int sweeps, final_sweeps, clamp_effort;
};

// task constants; every derived value is formed in double on the host the way Python forms it, then cast once
struct TaskParams {
  int N, skipframe, perturb, randomize, mocap_rows;
  int mocap_data_num;       // 3599, T:114
  float period;             // float32(3599*0.0005), T:116
  float cycle_dt;           // float32(0.0005), T:115
  float dt;                 // float32(sim dt), T:529
  float dt_policy;          // float32(dt*skipframe), T:540
  float time_gain;          // float32(5*dt_policy), T:541
  float pert_period;        // float32(8/dt_policy), T:495
  float max_len_m1;         // float32(max_episode_length-1), T:594 / VT:325
  float gate_len;           // max_episode_length - 8/dt_policy, T:489
  float death_cost, initial_height;
  float noise_std;          // float32(0.00016/3.0), T:528
  float dr_damping_base, dr_damping_lo, dr_damping_hi, dr_armature_lo, dr_armature_hi;
  int lfoot, rfoot, pelvis;
  uint64_t seed;
  // small device tables
  const float* kp;
  const float* kv;
  const float* action_high;
  const float* reset_dof_pos;   // clamp(initial_dof_pos, lower, upper), T:742
  const float* init_dof_pos;
  const float* armature_base;
  uint64_t* step_counter;       // device, Philox epoch; bumped once per step by the cross-env kernel
};

struct Sim {
  SimParams p;
  DevModel m;
  DyrosSimBuffers b;
  void* dev_blob = nullptr;  // one allocation holding every model table
  int device = 0;
};

struct Task {
  Sim* sim;
  TaskParams p;
  DyrosTaskBuffers b;
  DyrosNoiseInjection inj;
  void* dev_blob = nullptr;
};

// launchers (task_kernels.cu)
int launch_prologue(Task* t, const float* actions, cudaStream_t s);
int launch_substep_torque(Task* t, cudaStream_t s);
int launch_sensor_noise(Task* t, int substep, cudaStream_t s);
int launch_epilogue(Task* t, cudaStream_t s);
int launch_check_termination(Task* t, cudaStream_t s);
int launch_compute_reward(Task* t, cudaStream_t s);
int launch_crossenv(Task* t, bool compact, bool gate, bool bump, cudaStream_t s);
int launch_reset_idx(Task* t, const int64_t* env_ids, int count, cudaStream_t s);
int launch_compute_observations(Task* t, cudaStream_t s);
int launch_late_update(Task* t, cudaStream_t s);
int launch_post_fused(Task* t, cudaStream_t s);
// launchers (physics_kernels.cu)
int launch_simulate(Sim* sim, int apply_wrench, cudaStream_t s);
int launch_task_physics(Task* t, cudaStream_t s);  // skipframe x (PD + delay + substep + noise) in one launch
int launch_refresh_rigid_body_state(Sim* sim, cudaStream_t s);

}  // namespace dyros
