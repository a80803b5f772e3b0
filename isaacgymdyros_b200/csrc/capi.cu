// C ABI of libdyros_b200.so (include/dyros_b200.h): object lifetime, model-table upload, parameter
// derivation and the per-call launch sequences. No kernel lives here.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <new>
#include <string>
#include <vector>

#include "internal.h"

namespace dyros {

static thread_local std::string g_error;
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
}

// Host-side builder of one device blob: arrays are appended, then uploaded with a single copy.
struct Blob {
  std::vector<unsigned char> host;
  size_t add(const void* p, size_t bytes) {
    size_t off = (host.size() + 15) & ~size_t(15);
    host.resize(off + bytes);
    if (bytes) memcpy(host.data() + off, p, bytes);
    return off;
  }
  size_t add_f(const double* p, size_t n) {
    std::vector<float> f(n);
    for (size_t i = 0; i < n; ++i) f[i] = (float)p[i];
    return add(f.data(), n * sizeof(float));
  }
  size_t add_f32(const float* p, size_t n) { return add(p, n * sizeof(float)); }
  size_t add_i(const int* p, size_t n) { return add(p, n * sizeof(int)); }
};

template <class T>
static const T* at(void* base, size_t off) {
  return reinterpret_cast<const T*>(static_cast<unsigned char*>(base) + off);
}

#define REQUIRE(cond, ...)      \
  do {                          \
    if (!(cond)) {              \
      set_error(__VA_ARGS__);   \
      return 1;                 \
    }                           \
  } while (0)

static int build_sim(const DyrosSimDesc* d, const DyrosModelDesc* m, const DyrosSimBuffers* b, Sim** out) {
  REQUIRE(d && m && b && out, "dyros_sim_create: null argument");
  REQUIRE(d->num_envs > 0, "dyros_sim_create: num_envs must be positive (got %d)", d->num_envs);
  REQUIRE(d->substeps >= 1 && d->dt > 0, "dyros_sim_create: need dt > 0 and substeps >= 1");
  REQUIRE(m->num_links >= 1 && m->num_links <= DYROS_MAX_LINKS, "dyros_sim_create: num_links %d outside [1,%d]",
          m->num_links, DYROS_MAX_LINKS);
  REQUIRE(m->num_bodies >= m->num_links && m->num_bodies <= DYROS_MAX_BODIES, "dyros_sim_create: num_bodies %d",
          m->num_bodies);
  REQUIRE(m->num_dofs == m->num_links - 1, "dyros_sim_create: one revolute DOF per non-base link expected");
  REQUIRE(m->sched_slots >= 1 && m->sched, "dyros_sim_create: missing branch schedule");
  REQUIRE(b->root_states && b->dof_state && b->net_contact_force && b->dof_actuation_force && b->dof_damping &&
              b->dof_armature && b->body_mass_scale,
          "dyros_sim_create: a required device buffer is NULL");
  const int nl = m->num_links, nb = m->num_bodies, nd = m->num_dofs, np = m->num_points, nc = m->num_cyls;
  for (int l = 1; l < nl; ++l) {
    REQUIRE(m->link_parent[l] >= 0 && m->link_parent[l] < l, "dyros_sim_create: link %d parent %d not topological", l,
            m->link_parent[l]);
    REQUIRE(m->link_dof[l] >= 0 && m->link_dof[l] < nd, "dyros_sim_create: link %d dof index %d", l, m->link_dof[l]);
  }
  // every link exactly once in the schedule, after its parent
  {
    std::vector<int> slot(nl, -1);
    slot[0] = -1;
    for (int t = 0; t < m->sched_slots; ++t)
      for (int g = 0; g < DYROS_LANES; ++g) {
        int l = m->sched[t * DYROS_LANES + g];
        if (l < 0) continue;
        REQUIRE(l >= 1 && l < nl && slot[l] < 0, "dyros_sim_create: bad schedule entry %d", l);
        int p = m->link_parent[l];
        REQUIRE(p == 0 || (slot[p] >= 0 && slot[p] < t), "dyros_sim_create: schedule runs link %d before its parent", l);
        slot[l] = t;
      }
    for (int l = 1; l < nl; ++l) REQUIRE(slot[l] >= 0, "dyros_sim_create: link %d missing from the schedule", l);
  }

  Sim* sim = new (std::nothrow) Sim();
  REQUIRE(sim, "out of host memory");
  sim->device = d->device;
  SimParams& p = sim->p;
  p.N = d->num_envs;
  p.substeps = d->substeps;
  p.dt = (float)(d->dt / (double)d->substeps);
  for (int i = 0; i < 3; ++i) p.g[i] = d->gravity[i];
  p.contact_offset = d->contact_offset;
  p.max_depen_vel = d->max_depenetration_velocity;
  p.erp = d->contact_erp;
  p.mu = d->friction;
  p.pen_k = d->penalty_stiffness;
  p.pen_c = d->penalty_damping;
  p.pen_fmax = d->penalty_max_force;
  p.max_ang_vel = d->max_angular_velocity;
  p.sweeps = d->contact_sweeps;
  p.clamp_effort = d->clamp_effort;
  sim->b = *b;

  // ---- derived tables
  std::vector<int> child_start(nl + 1, 0), children;
  for (int l = 0; l < nl; ++l) {
    child_start[l] = (int)children.size();
    for (int c = 1; c < nl; ++c)
      if (m->link_parent[c] == l) children.push_back(c);
  }
  child_start[nl] = (int)children.size();
  std::vector<int> body_start(nl + 1, 0), bodies;
  for (int l = 0; l < nl; ++l) {
    body_start[l] = (int)bodies.size();
    for (int bb = 0; bb < nb; ++bb)
      if (m->body_link[bb] == l) bodies.push_back(bb);
  }
  body_start[nl] = (int)bodies.size();
  REQUIRE((int)bodies.size() == nb, "dyros_sim_create: body_link has entries outside [0,%d)", nl);
  std::vector<int> pt_start(nl + 1, 0), ppt_body;
  std::vector<float> ppt_pos, ppt_rad;
  for (int l = 0; l < nl; ++l) {
    pt_start[l] = (int)ppt_body.size();
    for (int i = 0; i < np; ++i)
      if (m->pt_link[i] == l && !(m->pt_solver && m->pt_solver[i])) {
        ppt_body.push_back(m->pt_body[i]);
        for (int k = 0; k < 3; ++k) ppt_pos.push_back((float)m->pt_pos[3 * i + k]);
        ppt_rad.push_back((float)m->pt_radius[i]);
      }
  }
  pt_start[nl] = (int)ppt_body.size();
  std::vector<int> cyl_start(nl + 1, 0), ccyl_body;
  std::vector<float> ccyl_center, ccyl_axis, ccyl_size;
  for (int l = 0; l < nl; ++l) {
    cyl_start[l] = (int)ccyl_body.size();
    for (int i = 0; i < nc; ++i)
      if (m->cyl_link[i] == l) {
        ccyl_body.push_back(m->cyl_body[i]);
        for (int k = 0; k < 3; ++k) ccyl_center.push_back((float)m->cyl_center[3 * i + k]);
        for (int k = 0; k < 3; ++k) ccyl_axis.push_back((float)m->cyl_axis[3 * i + k]);
        for (int k = 0; k < 2; ++k) ccyl_size.push_back((float)m->cyl_size[2 * i + k]);
      }
  }
  cyl_start[nl] = (int)ccyl_body.size();

  DevModel& dm = sim->m;
  memset(&dm, 0, sizeof(dm));
  dm.nl = nl; dm.nb = nb; dm.nd = nd; dm.np = (int)ppt_body.size(); dm.nc = (int)ccyl_body.size(); dm.T = m->sched_slots;
  // solver (foot) links, their chains and candidate points, in ascending link order
  for (int i = 0; i < np; ++i) {
    if (!(m->pt_solver && m->pt_solver[i])) continue;
    int l = m->pt_link[i], f = -1;
    for (int k = 0; k < dm.num_feet; ++k)
      if (dm.foot_link[k] == l) f = k;
    if (f < 0) {
      if (dm.num_feet >= MAX_FEET) {
        delete sim;
        set_error("dyros_sim_create: more than %d solver links", MAX_FEET);
        return 1;
      }
      f = dm.num_feet++;
      dm.foot_link[f] = l;
    }
    if (dm.foot_npts[f] >= MAX_SOLVER_PTS) {
      delete sim;
      set_error("dyros_sim_create: more than %d solver points on link %d", MAX_SOLVER_PTS, l);
      return 1;
    }
    int k = dm.foot_npts[f]++;
    dm.foot_pt_body[f][k] = m->pt_body[i];
    for (int c = 0; c < 3; ++c) dm.foot_pt_pos[f][k][c] = (float)m->pt_pos[3 * i + c];
    dm.foot_pt_radius[f][k] = (float)m->pt_radius[i];
  }
  if (dm.num_feet == 2 && dm.foot_link[0] > dm.foot_link[1]) {
    std::swap(dm.foot_link[0], dm.foot_link[1]);
    std::swap(dm.foot_npts[0], dm.foot_npts[1]);
    for (int k = 0; k < MAX_SOLVER_PTS; ++k) {
      std::swap(dm.foot_pt_body[0][k], dm.foot_pt_body[1][k]);
      std::swap(dm.foot_pt_radius[0][k], dm.foot_pt_radius[1][k]);
      for (int c = 0; c < 3; ++c) std::swap(dm.foot_pt_pos[0][k][c], dm.foot_pt_pos[1][k][c]);
    }
  }
  for (int f = 0; f < dm.num_feet; ++f) {
    std::vector<int> path;
    for (int l = dm.foot_link[f]; l > 0; l = m->link_parent[l]) path.push_back(l);
    if ((int)path.size() > MAX_CHAIN || path.empty()) {
      delete sim;
      set_error("dyros_sim_create: solver link %d is %zu joints from the base (max %d)", dm.foot_link[f], path.size(),
                MAX_CHAIN);
      return 1;
    }
    dm.chain_len[f] = (int)path.size();
    for (int k = 0; k < (int)path.size(); ++k) dm.chain[f][k] = path[path.size() - 1 - k];
  }
  if (dm.num_feet == 2) {  // the two chains must only share the base (block-Jacobi coupling goes through the base)
    for (int a = 0; a < dm.chain_len[0]; ++a)
      for (int c = 0; c < dm.chain_len[1]; ++c)
        if (dm.chain[0][a] == dm.chain[1][c]) {
          delete sim;
          set_error("dyros_sim_create: solver links %d and %d share link %d below the base", dm.foot_link[0],
                    dm.foot_link[1], dm.chain[0][a]);
          return 1;
        }
  }

  Blob bl;
  size_t o_parent = bl.add_i(m->link_parent, nl), o_dof = bl.add_i(m->link_dof, nl);
  size_t o_E = bl.add_f(m->link_E, nl * 9), o_r = bl.add_f(m->link_r, nl * 3), o_ax = bl.add_f(m->link_axis, nl * 3);
  size_t o_cs = bl.add_i(child_start.data(), nl + 1), o_ch = bl.add_i(children.data(), children.size());
  size_t o_bs = bl.add_i(body_start.data(), nl + 1), o_bd = bl.add_i(bodies.data(), nb);
  size_t o_bl = bl.add_i(m->body_link, nb), o_bp = bl.add_f(m->body_pos, nb * 3), o_br = bl.add_f(m->body_rot, nb * 9);
  size_t o_bi = bl.add_f(m->body_inertia, nb * 10);
  size_t o_lo = bl.add_f(m->dof_lower, nd), o_up = bl.add_f(m->dof_upper, nd), o_vl = bl.add_f(m->dof_vel_limit, nd);
  size_t o_ef = bl.add_f(m->dof_effort, nd);
  size_t o_ps = bl.add_i(pt_start.data(), nl + 1), o_pb = bl.add_i(ppt_body.data(), ppt_body.size());
  size_t o_pp = bl.add_f32(ppt_pos.data(), ppt_pos.size()), o_pr = bl.add_f32(ppt_rad.data(), ppt_rad.size());
  size_t o_ys = bl.add_i(cyl_start.data(), nl + 1), o_yb = bl.add_i(ccyl_body.data(), ccyl_body.size());
  size_t o_yc = bl.add_f32(ccyl_center.data(), ccyl_center.size()), o_ya = bl.add_f32(ccyl_axis.data(), ccyl_axis.size());
  size_t o_yz = bl.add_f32(ccyl_size.data(), ccyl_size.size());
  size_t o_sc = bl.add_i(m->sched, (size_t)m->sched_slots * DYROS_LANES);

  cudaError_t e = cudaSetDevice(d->device);
  if (e == cudaSuccess) e = cudaMalloc(&sim->dev_blob, bl.host.size());
  if (e == cudaSuccess) e = cudaMemcpy(sim->dev_blob, bl.host.data(), bl.host.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("dyros_sim_create: CUDA error while uploading the model: %s", cudaGetErrorString(e));
    if (sim->dev_blob) cudaFree(sim->dev_blob);
    delete sim;
    return 1;
  }
  void* base = sim->dev_blob;
  dm.link_parent = at<int>(base, o_parent); dm.link_dof = at<int>(base, o_dof);
  dm.link_E = at<float>(base, o_E); dm.link_r = at<float>(base, o_r); dm.link_axis = at<float>(base, o_ax);
  dm.link_child_start = at<int>(base, o_cs); dm.link_children = at<int>(base, o_ch);
  dm.link_body_start = at<int>(base, o_bs); dm.link_bodies = at<int>(base, o_bd);
  dm.body_link = at<int>(base, o_bl); dm.body_pos = at<float>(base, o_bp); dm.body_rot = at<float>(base, o_br);
  dm.body_inertia = at<float>(base, o_bi);
  dm.dof_lower = at<float>(base, o_lo); dm.dof_upper = at<float>(base, o_up); dm.dof_vel_limit = at<float>(base, o_vl);
  dm.dof_effort = at<float>(base, o_ef);
  dm.link_pt_start = at<int>(base, o_ps); dm.pt_body = at<int>(base, o_pb); dm.pt_pos = at<float>(base, o_pp);
  dm.pt_radius = at<float>(base, o_pr);
  dm.link_cyl_start = at<int>(base, o_ys); dm.cyl_body = at<int>(base, o_yb); dm.cyl_center = at<float>(base, o_yc);
  dm.cyl_axis = at<float>(base, o_ya); dm.cyl_size = at<float>(base, o_yz);
  dm.sched = at<int>(base, o_sc);

  int dev_sms = 0;
  if (cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, d->device) == cudaSuccess && dev_sms > 0)
    sim->sm_count = dev_sms;
  if (physics_configure(sim)) {
    cudaFree(sim->dev_blob);
    delete sim;
    return 1;
  }
  *out = sim;
  return 0;
}

static int build_task(Sim* sim, const DyrosTaskDesc* d, const DyrosTaskBuffers* b, Task** out) {
  REQUIRE(sim && d && b && out, "dyros_task_create: null argument");
  REQUIRE(sim->m.nd == ND && sim->m.nb == NB, "dyros_task_create: DyrosDynamicWalk needs %d DOF / %d bodies (model has %d / %d)",
          ND, NB, sim->m.nd, sim->m.nb);
  REQUIRE(d->skipframe >= 1 && d->skipframe <= 8, "dyros_task_create: skipframe %d", d->skipframe);
  REQUIRE(d->kp && d->kv && d->action_high && d->initial_dof_pos, "dyros_task_create: missing gain/limit tables");
  REQUIRE(d->mocap_rows >= 2 && b->mocap_data && b->obs_mean && b->obs_var, "dyros_task_create: missing shared tables");
  // every per-env pointer is required (names: DyrosTaskBuffers)
  const void* req[] = {b->obs_buf, b->rew_buf, b->reset_buf, b->timeout_buf, b->progress_buf, b->randomize_buf,
                       b->stacked_rewards, b->reset_env_ids, b->reset_env_ids32, b->reset_count, b->actions,
                       b->actions_pre, b->time, b->init_mocap_data_idx, b->mocap_data_idx, b->target_data_qpos,
                       b->target_data_force, b->action_torque, b->action_torque_pre, b->motor_constant_scale,
                       b->action_log, b->delay_idx, b->simul_len, b->qpos_noise, b->qvel_noise, b->qpos_pre,
                       b->qpos_bias, b->quat_bias, b->target_vel, b->pre_joint_velocity_states, b->contact_forces_pre,
                       b->total_mass, b->env_origins, b->epi_len, b->epi_len_log, b->contact_reward_sum,
                       b->contact_reward_mean, b->perturbation_count, b->pert_duration, b->pert_on, b->impulse,
                       b->magnitude, b->phase, b->perturb_timing, b->perturb_start, b->push_force, b->obs_history,
                       b->action_history, b->obs_hist_head, b->act_hist_head};
  for (size_t i = 0; i < sizeof(req) / sizeof(req[0]); ++i)
    REQUIRE(req[i], "dyros_task_create: DyrosTaskBuffers member #%zu is NULL", i);
  Task* t = new (std::nothrow) Task();
  REQUIRE(t, "out of host memory");
  t->sim = sim;
  t->b = *b;
  memset(&t->inj, 0, sizeof(t->inj));
  TaskParams& p = t->p;
  memset(&p, 0, sizeof(p));
  // Constants are formed in double exactly as the reference's Python forms them, then cast once.
  const double dt = (double)sim->p.dt * sim->p.substeps;  // gym.simulate advances sim.dt, T:118
  const double dt_sim = std::round(dt * 1e9) / 1e9;       // undo the float round trip of SimParams.dt (0.002)
  const double dt_policy = dt_sim * d->skipframe;         // T:120
  p.N = sim->p.N;
  p.skipframe = d->skipframe;
  p.perturb = d->perturb;
  p.randomize = d->randomize;
  p.mocap_rows = d->mocap_rows;
  p.mocap_data_num = d->mocap_rows - 1;                         // T:114
  p.cycle_dt = (float)0.0005;                                   // T:115
  p.period = (float)((double)p.mocap_data_num * 0.0005);        // T:116
  p.dt = (float)dt_sim;                                         // T:529
  p.dt_policy = (float)dt_policy;                               // T:540
  p.time_gain = (float)(5 * dt_policy);                         // T:541
  p.pert_period = (float)(8 / dt_policy);                       // T:495
  p.max_len_m1 = (float)((double)d->max_episode_length - 1);    // T:594, VT:325
  p.gate_len = (float)((double)d->max_episode_length - 8 / dt_policy);  // T:489
  p.death_cost = d->death_cost;
  p.initial_height = d->initial_height;
  p.noise_std = (float)(0.00016 / 3.0);                         // T:528
  p.dr_damping_base = d->dr_damping_base;
  p.dr_damping_lo = d->dr_damping_lo;
  p.dr_damping_hi = d->dr_damping_hi;
  p.dr_armature_lo = d->dr_armature_lo;
  p.dr_armature_hi = d->dr_armature_hi;
  p.lfoot = d->left_foot_body;
  p.rfoot = d->right_foot_body;
  p.pelvis = d->pelvis_body;
  p.seed = d->seed;
  REQUIRE(p.lfoot >= 0 && p.lfoot < NB && p.rfoot >= 0 && p.rfoot < NB && p.pelvis >= 0 && p.pelvis < NB,
          "dyros_task_create: body index out of range");

  std::vector<float> lower(ND), upper(ND), reset_pos(ND), arm(ND, 0.f);
  if (cudaMemcpy(lower.data(), sim->m.dof_lower, ND * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(upper.data(), sim->m.dof_upper, ND * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) {
    set_error("dyros_task_create: cannot read back dof limits");
    delete t;
    return 1;
  }
  for (int i = 0; i < ND; ++i)  // tensor_clamp(initial_dof_pos, lower, upper): max(min(x, upper), lower), TU:208, T:742
    reset_pos[i] = std::max(std::min(d->initial_dof_pos[i], upper[i]), lower[i]);
  if (d->dr_armature_base)
    for (int i = 0; i < ND; ++i) arm[i] = (float)d->dr_armature_base[i];
  Blob bl;
  size_t o_kp = bl.add_f32(d->kp, ND), o_kv = bl.add_f32(d->kv, ND), o_ah = bl.add_f32(d->action_high, ND);
  size_t o_rp = bl.add_f32(reset_pos.data(), ND), o_ip = bl.add_f32(d->initial_dof_pos, ND), o_ar = bl.add_f32(arm.data(), ND);
  unsigned long long zero = 0;
  size_t o_ct = bl.add(&zero, sizeof(zero));
  cudaError_t e = cudaMalloc(&t->dev_blob, bl.host.size());
  if (e == cudaSuccess) e = cudaMemcpy(t->dev_blob, bl.host.data(), bl.host.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("dyros_task_create: CUDA error: %s", cudaGetErrorString(e));
    if (t->dev_blob) cudaFree(t->dev_blob);
    delete t;
    return 1;
  }
  void* base = t->dev_blob;
  p.kp = at<float>(base, o_kp); p.kv = at<float>(base, o_kv); p.action_high = at<float>(base, o_ah);
  p.reset_dof_pos = at<float>(base, o_rp); p.init_dof_pos = at<float>(base, o_ip); p.armature_base = at<float>(base, o_ar);
  p.step_counter = const_cast<uint64_t*>(at<uint64_t>(base, o_ct));
  *out = t;
  return 0;
}

}  // namespace dyros

using namespace dyros;

extern "C" {

const char* dyros_last_error(void) { return g_error.c_str(); }
int dyros_abi_version(void) { return DYROS_ABI_VERSION; }

int dyros_sim_create(const DyrosSimDesc* desc, const DyrosModelDesc* model, const DyrosSimBuffers* buf, DyrosSim** out) {
  Sim* s = nullptr;
  int rc = build_sim(desc, model, buf, &s);
  if (rc == 0) *out = reinterpret_cast<DyrosSim*>(s);
  return rc;
}
int dyros_sim_destroy(DyrosSim* sim) {
  Sim* s = reinterpret_cast<Sim*>(sim);
  if (!s) return 0;
  if (s->dev_blob) cudaFree(s->dev_blob);
  delete s;
  return 0;
}
#define SIM_OR_FAIL(fn)                           \
  Sim* s = reinterpret_cast<Sim*>(sim);           \
  if (!s) {                                       \
    set_error(fn ": sim is NULL");                \
    return 1;                                     \
  }
#define TASK_OR_FAIL(fn)                          \
  Task* t = reinterpret_cast<Task*>(task);        \
  if (!t) {                                       \
    set_error(fn ": task is NULL");               \
    return 1;                                     \
  }

int dyros_simulate(DyrosSim* sim, int apply_wrench, void* stream) {
  SIM_OR_FAIL("dyros_simulate");
  if (apply_wrench && !(s->b.rb_force && s->b.rb_torque)) {
    set_error("dyros_simulate: apply_wrench set but rb_force / rb_torque buffers are NULL");
    return 1;
  }
  return launch_simulate(s, apply_wrench, nullptr, (cudaStream_t)stream);
}
int dyros_refresh_rigid_body_state(DyrosSim* sim, void* stream) {
  SIM_OR_FAIL("dyros_refresh_rigid_body_state");
  if (!s->b.rigid_body_state) {
    set_error("dyros_refresh_rigid_body_state: rigid_body_state buffer is NULL");
    return 1;
  }
  return launch_refresh_rigid_body_state(s, (cudaStream_t)stream);
}
int dyros_set_state_indexed(DyrosSim* sim, const int32_t* env_ids, int count, void* stream) {
  SIM_OR_FAIL("dyros_set_state_indexed");
  (void)stream;
  if (count < 0 || count > s->p.N || (count > 0 && !env_ids)) {
    set_error("dyros_set_state_indexed: count %d outside [0,%d] or NULL ids", count, s->p.N);
    return 1;
  }
  return 0;  // buffers are the live state (immediate CPU-pipeline semantics)
}

int dyros_task_create(DyrosSim* sim, const DyrosTaskDesc* desc, const DyrosTaskBuffers* buf, DyrosTask** out) {
  Task* t = nullptr;
  int rc = build_task(reinterpret_cast<Sim*>(sim), desc, buf, &t);
  if (rc == 0) *out = reinterpret_cast<DyrosTask*>(t);
  return rc;
}
int dyros_task_destroy(DyrosTask* task) {
  Task* t = reinterpret_cast<Task*>(task);
  if (!t) return 0;
  if (t->dev_blob) cudaFree(t->dev_blob);
  delete t;
  return 0;
}
int dyros_task_set_noise_injection(DyrosTask* task, const DyrosNoiseInjection* inj) {
  TASK_OR_FAIL("dyros_task_set_noise_injection");
  if (inj) t->inj = *inj;
  else memset(&t->inj, 0, sizeof(t->inj));
  if ((t->inj.pert_i != nullptr) != (t->inj.pert_f != nullptr)) {
    set_error("dyros_task_set_noise_injection: pert_i and pert_f must be given together");
    memset(&t->inj, 0, sizeof(t->inj));
    return 1;
  }
  return 0;
}
int dyros_task_prologue(DyrosTask* task, const float* actions, void* stream) {
  TASK_OR_FAIL("dyros_task_prologue");
  if (!actions) {
    set_error("dyros_task_prologue: actions is NULL");
    return 1;
  }
  return launch_prologue(t, actions, (cudaStream_t)stream);
}
int dyros_task_substep_torque(DyrosTask* task, void* stream) {
  TASK_OR_FAIL("dyros_task_substep_torque");
  return launch_substep_torque(t, (cudaStream_t)stream);
}
int dyros_task_sensor_noise(DyrosTask* task, int substep, void* stream) {
  TASK_OR_FAIL("dyros_task_sensor_noise");
  if (substep < 0 || substep >= t->p.skipframe) {
    set_error("dyros_task_sensor_noise: substep %d outside [0,%d)", substep, t->p.skipframe);
    return 1;
  }
  return launch_sensor_noise(t, substep, (cudaStream_t)stream);
}
int dyros_task_epilogue(DyrosTask* task, void* stream) {
  TASK_OR_FAIL("dyros_task_epilogue");
  return launch_epilogue(t, (cudaStream_t)stream);
}
int dyros_task_check_termination(DyrosTask* task, void* stream) {
  TASK_OR_FAIL("dyros_task_check_termination");
  return launch_check_termination(t, (cudaStream_t)stream);
}
int dyros_task_compute_reward(DyrosTask* task, void* stream) {
  TASK_OR_FAIL("dyros_task_compute_reward");
  return launch_compute_reward(t, (cudaStream_t)stream);
}
int dyros_task_compact_resets(DyrosTask* task, void* stream) {
  TASK_OR_FAIL("dyros_task_compact_resets");
  return launch_crossenv(t, true, false, false, (cudaStream_t)stream);
}
int dyros_task_reset_idx(DyrosTask* task, const int64_t* env_ids, int count, void* stream) {
  TASK_OR_FAIL("dyros_task_reset_idx");
  if (env_ids && (count < 0 || count > t->p.N)) {
    set_error("dyros_task_reset_idx: count %d outside [0,%d]", count, t->p.N);
    return 1;
  }
  return launch_reset_idx(t, env_ids, env_ids ? count : -1, (cudaStream_t)stream);
}
int dyros_task_compute_observations(DyrosTask* task, void* stream) {
  TASK_OR_FAIL("dyros_task_compute_observations");
  return launch_compute_observations(t, (cudaStream_t)stream);
}
int dyros_task_late_update(DyrosTask* task, void* stream) {
  TASK_OR_FAIL("dyros_task_late_update");
  return launch_late_update(t, (cudaStream_t)stream);
}
int dyros_task_end_step(DyrosTask* task, void* stream) {
  TASK_OR_FAIL("dyros_task_end_step");
  return launch_crossenv(t, false, true, true, (cudaStream_t)stream);
}

int dyros_task_step(DyrosTask* task, const float* actions, void* stream) {
  TASK_OR_FAIL("dyros_task_step");
  if (!actions) {
    set_error("dyros_task_step: actions is NULL");
    return 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (launch_prologue(t, actions, st)) return 1;
  for (int k = 0; k < t->p.skipframe; ++k) {
    if (launch_substep_torque(t, st)) return 1;
    // the pelvis push acts on the first sub-step only: applied once before the loop, T:502 vs T:504
    if (launch_simulate(t->sim, 0, k == 0 ? t->b.push_force : nullptr, st)) return 1;
    if (launch_sensor_noise(t, k, st)) return 1;
  }
  if (launch_post_fused(t, st)) return 1;
  return launch_crossenv(t, true, true, true, st);
}
int dyros_task_step_launches(DyrosTask* task) {
  TASK_OR_FAIL("dyros_task_step_launches");
  return 1 + 3 * t->p.skipframe + 2;
}

}  // extern "C"
