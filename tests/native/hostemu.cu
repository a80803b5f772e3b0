// TEST INFRASTRUCTURE ONLY. CPU emulation of the CUDA physics kernel's lane program: the very same source
// (isaacgymdyros_b200/csrc/physics_core.cuh) compiled for the host, DYROS_LANES threads per env meeting at a
// barrier where the GPU lanes meet at __syncwarp(). Lets the CPU test suite compare the kernel's O(n)
// recursions with the dense fp64 oracle (oracle/physics_oracle.py) without a GPU. Never linked into the product.
#include <barrier>
#include <thread>
#include <vector>

#include "host_model.h"
#include "physics_core.cuh"

namespace dyros {
void set_error(const char*, ...) {}
}  // namespace dyros
using namespace dyros;

struct BarrierSync {
  std::barrier<>* b;
  void operator()() const { b->arrive_and_wait(); }
};

extern "C" int dyros_hostemu_simulate(const DyrosSimDesc* d, const DyrosModelDesc* md, float* root, float* dof_state,
                                      const float* tau, const float* damping, const float* armature,
                                      const float* mass_scale, float* contact, const float* push, const float* rb_force,
                                      const float* rb_torque, char* err, int errlen) {
  Blob bl;
  DevModel m;
  ModelOffsets off;
  std::string e = build_model_tables(md, bl, m, off);
  if (!e.empty()) {
    snprintf(err, errlen, "%s", e.c_str());
    return 1;
  }
  resolve_model(m, off, bl.host.data());
  SimParams p;
  fill_sim_params(d, p);
  const int es = env_scratch_floats(m.nl);
  std::vector<real> scratch(es);
  std::barrier<> bar(DYROS_LANES);
  for (int env = 0; env < p.N; ++env) {
    EnvIO io;
    io.root = root + (size_t)env * 13;
    io.dof_state = dof_state + (size_t)env * m.nd * 2;
    io.tau = tau + (size_t)env * m.nd;
    io.damping = damping + (size_t)env * m.nd;
    io.armature = armature + (size_t)env * m.nd;
    io.mass_scale = mass_scale + (size_t)env * m.nb;
    io.contact = contact + (size_t)env * m.nb * 3;
    io.push = push ? push + (size_t)env * 3 : nullptr;
    io.rb_force = rb_force ? rb_force + (size_t)env * m.nb * 3 : nullptr;
    io.rb_torque = rb_torque ? rb_torque + (size_t)env * m.nb * 3 : nullptr;
    io.live = true;
    for (int s = 0; s < p.substeps; ++s) {
      std::vector<std::thread> th;
      for (int g = 0; g < DYROS_LANES; ++g)
        th.emplace_back([&, g]() {
          BarrierSync sync{&bar};
          env_substep(io, scratch.data(), reinterpret_cast<const float*>(bl.host.data()), m, p, g, sync);
        });
      for (auto& t : th) t.join();
      io.push = nullptr;
      io.rb_force = nullptr;
      io.rb_torque = nullptr;
    }
  }
  return 0;
}

extern "C" int dyros_hostemu_layout(const DyrosModelDesc* md, int* hot_bytes, int* env_scratch_bytes) {
  Blob bl;
  DevModel m;
  ModelOffsets off;
  if (!build_model_tables(md, bl, m, off).empty()) return 1;
  *hot_bytes = m.hot_bytes;
  *env_scratch_bytes = env_scratch_floats(m.nl) * (int)sizeof(real);
  return 0;
}
