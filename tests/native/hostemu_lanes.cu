// TEST INFRASTRUCTURE ONLY. CPU emulation of the multi-lane physics program (isaacgymdyros_b200/csrc/physics_lanes.cuh):
// the very same source compiled for the host with every lane-varying value emulated as 8 lanes (lane_emu.h), one host
// thread per role with atomic stage flags where the GPU uses release/acquire on shared memory. Lets the CPU test suite
// compare the kernel's recursions, its lane mapping (shuffles, per-lane gathers, masks) and its dataflow
// synchronisation with the dense fp64 oracle (oracle/physics_oracle.py) without a GPU. Never linked into the product.
#include "lane_emu.h"

#include <barrier>
#include <thread>
#include <vector>

#include "host_model.h"
#include "physics_lanes.cuh"

namespace dyros {
void set_error(const char*, ...) {}
}  // namespace dyros
using namespace dyros;

struct HostLaneSync {
  void mark(int) const {}
  void signal(int* f, int v) const { __atomic_store_n(f, v, __ATOMIC_RELEASE); }
  void wait(const int* f, int v) const {
    while (__atomic_load_n(f, __ATOMIC_ACQUIRE) < v) std::this_thread::yield();
  }
  void wait_io(const int* f, int v) const { wait(f, v); }
};

extern "C" int dyros_hostemu_lanes_simulate(const DyrosSimDesc* d, const DyrosModelDesc* md, float* root, float* dof_state,
                                            const float* tau, const float* damping, const float* armature,
                                            const float* mass_scale, float* contact, const float* push,
                                            const float* rb_force, const float* rb_torque, const float* friction,
                                            char* err, int errlen) {
  Blob bl;
  DevModel m;
  ModelOffsets off;
  std::string e = build_model_tables(md, bl, m, off);
  if (!e.empty()) {
    snprintf(err, errlen, "%s", e.c_str());
    return 1;
  }
  // 16-byte aligned copy of the tables (the program reads records with 128-bit loads)
  std::vector<float> blob_store(bl.host.size() / 4 + 8);
  float* blob = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(blob_store.data()) + 15) & ~uintptr_t(15));
  memcpy(blob, bl.host.data(), bl.host.size());
  resolve_model(m, off, blob);
  SimParams p;
  fill_sim_params(d, p);
  const float* hot = blob;
  const int es = ln::env_scratch_floats(m.nl, m.nb);
  std::vector<float> store(es + 8);
  float* sm = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(store.data()) + 15) & ~uintptr_t(15));
  for (int env = 0; env < p.N; ++env) {
    float* X = sm + m.nl * ln::LB;
    float* croot = root + (size_t)env * 13;
    float* cdof = dof_state + (size_t)env * m.nd * 2;
    ln::EnvIO io;
    io.contact = contact + (size_t)env * m.nb * 3;
    io.rb_force = rb_force ? rb_force + (size_t)env * m.nb * 3 : nullptr;
    io.rb_torque = rb_torque ? rb_torque + (size_t)env * m.nb * 3 : nullptr;
    io.push = push != nullptr;
    io.link_pose = nullptr;
    io.live = true;
    std::vector<int> flags(ln::QF_COUNT, 0);
    for (int s = 0; s < p.substeps; ++s) {
      // stage the inputs (the CUDA kernels do this with coalesced slab copies)
      for (int i = 0; i < es; ++i) sm[i] = 0.f;
      for (int b = 0; b < m.nb; ++b) X[ln::X_MASS + b] = mass_scale[(size_t)env * m.nb + b];
      for (int k = 0; k < 13; ++k) X[ln::X_ROOT + k] = croot[k];
      for (int k = 0; k < 3; ++k) X[ln::X_PUSH + k] = push ? push[(size_t)env * 3 + k] : 0.f;
      X[ln::X_MU] = friction ? friction[env] : p.mu;
      for (int k = 0; k < 3 * m.nb; ++k) io.contact[k] = 0.f;  // net contact force of THIS sub-step only
      for (int i = 1; i < m.nl; ++i) {
        const int dd = m.link_dof[i];
        float* L = sm + i * ln::LB;
        L[ln::B_Q] = cdof[2 * dd];
        L[ln::B_QD] = cdof[2 * dd + 1];
        L[ln::B_SC + 0] = tau[(size_t)env * m.nd + dd];
        L[ln::B_SC + 1] = damping[(size_t)env * m.nd + dd];
        L[ln::B_SC + 2] = armature[(size_t)env * m.nd + dd];
      }
      ln::EnvIO mine = io;  // applied wrenches act over the whole simulate() call, i.e. all of its sub-steps
      std::vector<std::thread> th;
      for (int role = 0; role < DYROS_LANES; ++role)
        th.emplace_back([&, role]() {
          HostLaneSync sync;
          const Ln g = make_ln(0);
          ln::env_substep_lanes(mine, sm, flags.data(), nullptr, s, hot, m, p, role, sync, g, false);
        });
      for (auto& t : th) t.join();
      for (int k = 0; k < 13; ++k) croot[k] = X[ln::X_ROOT + k];
      for (int i = 1; i < m.nl; ++i) {
        const int dd = m.link_dof[i];
        cdof[2 * dd] = sm[i * ln::LB + ln::B_Q];
        cdof[2 * dd + 1] = sm[i * ln::LB + ln::B_QD];
      }
    }
  }
  return 0;
}

// Dynamic shared memory of a physics CTA holding `epb` envs (mirrors phys_smem_bytes in physics_kernels.cu).
extern "C" long dyros_hostemu_lanes_cta_smem_bytes(const DyrosModelDesc* md, int epb) {
  Blob bl;
  DevModel m;
  ModelOffsets off;
  if (!build_model_tables(md, bl, m, off).empty()) return -1;
  const int nquad = (epb + 3) / 4;
  return (long)m.hot_bytes + ((long)ln::IOF_COUNT + (long)nquad * ln::QF_COUNT + (long)epb * ln::env_scratch_floats(m.nl, m.nb)) * 4L;
}
