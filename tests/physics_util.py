"""Shared helpers of the physics tests: oracle parameters from a CoreConfig, random states, and the CPU
emulation of the kernel's lane program (tests/native/hostemu.cu)."""
import ctypes as C
import os
import subprocess

import numpy as np

from isaacgymdyros_b200 import native
from isaacgymdyros_b200.core import ARMATURE, INIT_DOF_POS, CoreConfig, make_model_desc, make_sim_desc
from oracle.physics_oracle import PhysParams

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "isaacgymdyros_b200", "csrc")


def oracle_params(cfg: CoreConfig) -> PhysParams:
    return PhysParams(dt=cfg.dt / cfg.substeps, gravity=cfg.gravity, contact_offset=cfg.contact_offset,
                      max_depen_vel=cfg.max_depenetration_velocity, erp=cfg.contact_erp, mu=cfg.friction,
                      pen_k=cfg.penalty_stiffness, pen_c=cfg.penalty_damping, pen_fmax=cfg.penalty_max_force,
                      max_ang_vel=cfg.max_angular_velocity,
                      sweeps=cfg.num_position_iterations + cfg.num_velocity_iterations,
                      clamp_effort=cfg.clamp_effort, vel_limit=cfg.dof_vel_limit)


def random_states(N, rng, tables, kind="mixed"):
    """float32-representable random states. kind: 'air' (no contact), 'stand' (reset pose near the ground),
    'mixed' (airborne, standing and tilted/fallen envs)."""
    root = np.zeros((N, 13))
    root[:, 6] = 1.0
    q = np.tile(np.array(INIT_DOF_POS), (N, 1))
    qd = np.zeros((N, 33))
    mode = {"air": np.zeros(N, int), "stand": np.ones(N, int)}.get(kind)
    if mode is None:
        mode = rng.integers(0, 3, N)
    for n in range(N):
        if mode[n] == 0:  # airborne, arbitrary orientation and speed
            root[n, 2] = 2.0 + rng.uniform(0, 1)
            quat = rng.normal(0, 1, 4)
            root[n, 3:7] = quat / np.linalg.norm(quat)
            root[n, 7:13] = rng.normal(0, 0.5, 6)
            q[n] += rng.normal(0, 0.2, 33)
            qd[n] = rng.normal(0, 0.8, 33)
        elif mode[n] == 1:  # standing: soles within a few mm of the plane, small motion
            root[n, 2] = 0.93 + rng.uniform(-0.004, 0.002)
            quat = np.array([0, 0, 0, 1.0]) + np.append(rng.normal(0, 0.004, 3), 0)
            root[n, 3:7] = quat / np.linalg.norm(quat)
            root[n, 7:13] = rng.normal(0, 0.05, 6)
            q[n] += rng.normal(0, 0.003, 33)
            qd[n] = rng.normal(0, 0.1, 33)
        else:  # low and tilted: several links touch the ground (penalty contacts)
            root[n, 2] = rng.uniform(0.2, 0.8)
            quat = np.array([0, 0, 0, 1.0]) + np.append(rng.normal(0, 0.4, 3), 0)
            root[n, 3:7] = quat / np.linalg.norm(quat)
            root[n, 7:13] = rng.normal(0, 0.3, 6)
            q[n] += rng.normal(0, 0.3, 33)
            qd[n] = rng.normal(0, 0.5, 33)
    q = np.clip(q, tables.dof_lower + 0.05, tables.dof_upper - 0.05)
    f = lambda a: a.astype(np.float32).astype(np.float64)
    tau = rng.normal(0, 1, (N, 33)) * np.array(ARMATURE) * 40
    damping = 0.1 + rng.uniform(0, 2.9, (N, 33))
    armature = np.array(ARMATURE) * rng.uniform(0.8, 1.2, (N, 33))
    mass_scale = rng.uniform(0.8, 1.2, (N, 38))
    return dict(root=f(root), q=f(q), qd=f(qd), tau=f(tau), damping=f(damping), armature=f(armature),
                mass_scale=f(mass_scale))


_EMU = None


def hostemu():
    global _EMU
    if _EMU is None:
        so = os.path.join(HERE, "native", "libdyros_hostemu.so")
        src = os.path.join(HERE, "native", "hostemu.cu")
        deps = [src] + [os.path.join(CSRC, f) for f in ("physics_roles.cuh", "phys_math.cuh", "host_model.h", "internal.h")]
        if not os.path.isfile(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["nvcc", "-O2", "-std=c++20", "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "20014",
                                   "-Wno-deprecated-gpu-targets", "-I", CSRC, "-o", so, src])
        _EMU = C.CDLL(so)
        _EMU.dyros_hostemu_simulate.restype = C.c_int
    return _EMU


_EMU_LANES = {}


def hostemu_lanes(scalar="float"):
    """CPU build of the multi-lane program (csrc/physics_lanes.cuh through tests/native/lane_emu.h); scalar = "double"
    runs the same program in float64."""
    if scalar not in _EMU_LANES:
        so = os.path.join(HERE, "native", f"libdyros_hostemu_lanes_{scalar}.so")
        src = os.path.join(HERE, "native", "hostemu_lanes.cu")
        deps = [src, os.path.join(HERE, "native", "lane_emu.h")] + [
            os.path.join(CSRC, f) for f in ("physics_lanes.cuh", "lanes.cuh", "phys_math.cuh", "host_model.h", "internal.h")]
        if not os.path.isfile(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["nvcc", "-O1", "-std=c++20", "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "20014",
                                   "-Wno-deprecated-gpu-targets", f"-DEMU_SCALAR={scalar}", "-I", CSRC, "-o", so, src])
        lib = C.CDLL(so)
        lib.dyros_hostemu_lanes_simulate.restype = C.c_int
        lib.dyros_hostemu_lanes_cta_smem_bytes.restype = C.c_long
        _EMU_LANES[scalar] = lib
    return _EMU_LANES[scalar]


def emulate_substep_lanes(tables, cfg, st, push=None, rb_force=None, rb_torque=None, friction=None, scalar="float"):
    """Runs the multi-lane program of the CUDA kernel on the CPU. Returns root', q', qd', contact like the oracle."""
    return emulate_substep(tables, cfg, st, push, rb_force, rb_torque, friction,
                           fn=hostemu_lanes(scalar).dyros_hostemu_lanes_simulate)


def emulate_substep(tables, cfg, st, push=None, rb_force=None, rb_torque=None, friction=None, fn=None):
    """Runs the kernel's lane program on the CPU (float32). Returns root', q', qd', contact like the oracle."""
    N = st["root"].shape[0]
    md, keep = make_model_desc(tables, cfg)
    sd = make_sim_desc(cfg, N)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    root = f32(st["root"])
    dof = f32(np.stack([st["q"], st["qd"]], -1))
    tau, damp, arm, ms = f32(st["tau"]), f32(st["damping"]), f32(st["armature"]), f32(st["mass_scale"])
    contact = np.zeros((N, tables.num_bodies, 3), np.float32)
    P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    pf = f32(push) if push is not None else None
    rf = f32(rb_force) if rb_force is not None else None
    rt = f32(rb_torque) if rb_torque is not None else None
    err = C.create_string_buffer(256)
    mu = f32(friction) if friction is not None else None
    rc = (fn or hostemu().dyros_hostemu_simulate)(C.byref(sd), C.byref(md), P(root), P(dof), P(tau), P(damp), P(arm), P(ms),
                                          P(contact), P(pf), P(rf), P(rt), P(mu), err, 256)
    assert rc == 0, err.value
    return root.astype(np.float64), dof[:, :, 0].astype(np.float64), dof[:, :, 1].astype(np.float64), contact.astype(np.float64)
