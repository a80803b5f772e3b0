"""CPU check of the multi-lane physics program the CUDA kernels run (csrc/physics_lanes.cuh: 8 lanes per env, 6x6
inertias distributed by columns, lane = link phases, width-8 shuffles) against the dense fp64 oracle. The program is
compiled for the host with every lane-varying value emulated as 8 lanes (tests/native/lane_emu.h), one host thread per
role: this exercises the lane mapping (shuffles, per-lane gathers, masks, replicated-store checks) and the flag
protocol, not just the formulas. Same states and tolerances as tests/test_physics_emulation.py."""
import dataclasses

import numpy as np
import pytest

from isaacgymdyros_b200.core import CoreConfig
from oracle.physics_oracle import PhysicsOracle
from tests.golden_util import load_assets
from tests.physics_util import emulate_substep, emulate_substep_lanes, oracle_params, random_states
from tests.test_physics_emulation import SLIDING_SLACK, compare, sliding_states


@pytest.mark.parametrize("scalar", ["float", "double"])
@pytest.mark.parametrize("kind,seed", [("air", 0), ("stand", 1), ("mixed", 2)])
def test_lanes_program_matches_dense_oracle(kind, seed, scalar):
    tables = load_assets()[0]
    cfg = CoreConfig()
    o = PhysicsOracle(tables, oracle_params(cfg))
    rng = np.random.default_rng(seed)
    st = random_states(12, rng, tables, kind)
    push = rng.normal(0, 300, (12, 3))
    push[:, 2] = 0
    want = o.substep(st["root"], st["q"], st["qd"], st["tau"], st["damping"], st["armature"], st["mass_scale"], push=push)
    got = emulate_substep_lanes(tables, cfg, st, push=push, scalar=scalar)
    compare(st, got, want, ctx=f"{kind}/{scalar}: ")


def test_lanes_program_agrees_with_single_lane_program():
    """Two mappings of the same formulas: float32 results differ by rounding only."""
    tables = load_assets()[0]
    cfg = CoreConfig()
    rng = np.random.default_rng(4)
    st = random_states(16, rng, tables, "mixed")
    a = emulate_substep(tables, cfg, st)
    b = emulate_substep_lanes(tables, cfg, st)
    assert np.abs(a[1] - b[1]).max() < 1e-6 and np.abs(a[2] - b[2]).max() < 2e-3
    assert np.abs(a[0][:, :7] - b[0][:, :7]).max() < 1e-6
    assert np.abs(a[3] - b[3]).max() <= 1e-3 * max(1.0, np.abs(a[3]).max())


@pytest.mark.parametrize("kind", ["stand", "mixed"])
def test_lanes_program_per_env_friction(kind):
    tables = load_assets()[0]
    cfg = CoreConfig()
    N = 12
    rng = np.random.default_rng(21)
    st = sliding_states(N, rng, tables, kind)
    mu = rng.uniform(0.2, 1.3, N).astype(np.float32).astype(np.float64)
    o = PhysicsOracle(tables, dataclasses.replace(oracle_params(cfg), mu=mu))
    want = o.substep(st["root"], st["q"], st["qd"], st["tau"], st["damping"], st["armature"], st["mass_scale"])
    got = emulate_substep_lanes(tables, cfg, st, friction=mu)
    compare(st, got, want, ctx=f"{kind}: ", slack=SLIDING_SLACK)


def test_lanes_program_body_wrench_and_effort_clamp():
    tables = load_assets()[0]
    cfg = CoreConfig(clamp_effort=True, gravity=(0.0, 0.0, 0.0))
    o = PhysicsOracle(tables, oracle_params(cfg))
    rng = np.random.default_rng(5)
    st = random_states(4, rng, tables, "air")
    st["tau"] = (st["tau"] * 50).astype(np.float32).astype(np.float64)
    F, T = rng.normal(0, 50, (4, 38, 3)), rng.normal(0, 5, (4, 38, 3))
    want = o.substep(st["root"], st["q"], st["qd"], st["tau"], st["damping"], st["armature"], st["mass_scale"],
                     rb_force=F, rb_torque=T)
    got = emulate_substep_lanes(tables, cfg, st, rb_force=F, rb_torque=T)
    compare(st, got, want)


def test_lanes_program_short_rollout_stays_with_oracle():
    tables = load_assets()[0]
    cfg = CoreConfig()
    o = PhysicsOracle(tables, oracle_params(cfg))
    rng = np.random.default_rng(9)
    st = random_states(6, rng, tables, "stand")
    a, b = dict(st), dict(st)
    for _ in range(5):
        r, q, qd, c, _d = o.substep(a["root"], a["q"], a["qd"], a["tau"], a["damping"], a["armature"], a["mass_scale"])
        a.update(root=r, q=q, qd=qd)
        r2, q2, qd2, c2 = emulate_substep_lanes(tables, cfg, b)
        b.update(root=r2, q=q2, qd=qd2)
    assert np.abs(a["q"] - b["q"]).max() < 2e-5 and np.abs(a["qd"] - b["qd"]).max() < 5e-3
    assert np.abs(a["root"][:, :7] - b["root"][:, :7]).max() < 2e-5
    assert np.abs(c - c2).max() < 2e-3 * np.abs(c).max()


def test_tocabi_cta_holds_28_envs_in_the_lane_layout():
    """4096 envs on 148 SMs run as ONE wave only if a CTA (one per SM, 227 KiB of shared memory) holds 28 envs."""
    import ctypes as C
    from isaacgymdyros_b200.core import make_model_desc
    from tests.physics_util import hostemu_lanes
    md, keep = make_model_desc(load_assets()[0], CoreConfig())
    need = hostemu_lanes().dyros_hostemu_lanes_cta_smem_bytes(C.byref(md), 28)
    assert 0 < need <= 227 * 1024 - 1024, need  # 1 KiB stays free for static shared memory
