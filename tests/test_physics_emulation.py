"""CPU check of the CUDA physics kernel's algorithm: the kernel's role program (physics_roles.cuh, compiled for
the host, float32, O(n) articulated-body recursions) against the dense fp64 oracle (joint-space mass matrix).
Two independent formulations of the same model must agree to float32 round-off on single-step state deltas."""
import numpy as np
import pytest

from isaacgymdyros_b200.core import CoreConfig
from oracle.physics_oracle import PhysicsOracle
from tests.golden_util import load_assets
from tests.physics_util import emulate_substep, oracle_params, random_states

# per-quantity tolerance on single-step results: rtol * (largest single-step change of that quantity in the batch) + atol
TOL = {"root_pos": (1e-4, 2e-6), "root_quat": (1e-4, 2e-6), "root_vel": (1e-4, 1e-4), "q": (1e-4, 2e-6),
       "qd": (1e-4, 1e-4), "contact": (1e-4, 0.5)}


def compare(before, got, want, ctx="", slack=1.0):
    r0, q0, qd0 = before["root"], before["q"], before["qd"]
    rg, qg, qdg, cg = got
    rw, qw, qdw, cw = want[:4]
    items = {"root_pos": (rg[:, :3], rw[:, :3], r0[:, :3]), "root_quat": (rg[:, 3:7], rw[:, 3:7], r0[:, 3:7]),
             "root_vel": (rg[:, 7:], rw[:, 7:], r0[:, 7:]), "q": (qg, qw, q0), "qd": (qdg, qdw, qd0),
             "contact": (cg, cw, np.zeros_like(cw))}
    for k, (g, w, b) in items.items():
        rtol, atol = TOL[k]
        scale = np.abs(w - b).max()
        err = np.abs(g - w).max()
        assert err <= slack * (rtol * scale + atol), f"{ctx}{k}: err {err:.3e} > {slack}*({rtol}*{scale:.3e}+{atol})"


@pytest.mark.parametrize("kind,seed", [("air", 0), ("stand", 1), ("mixed", 2)])
def test_lane_program_matches_dense_oracle(kind, seed):
    tables = load_assets()[0]
    cfg = CoreConfig()
    o = PhysicsOracle(tables, oracle_params(cfg))
    rng = np.random.default_rng(seed)
    st = random_states(12, rng, tables, kind)
    push = rng.normal(0, 300, (12, 3))
    push[:, 2] = 0
    want = o.substep(st["root"], st["q"], st["qd"], st["tau"], st["damping"], st["armature"], st["mass_scale"], push=push)
    got = emulate_substep(tables, cfg, st, push=push)
    compare(st, got, want, ctx=f"{kind}: ")
    if kind == "stand":
        assert (want[3][:, [8, 16], 2] > 0).any(), "standing batch should load the soles"
    if kind == "mixed":
        non_feet = [b for b in range(38) if b not in (8, 16)]
        assert (np.abs(want[3][:, non_feet]).sum(-1) > 1).any(), "mixed batch should exercise penalty contacts"


SLIDING_SLACK = 3.0


def sliding_states(N, rng, tables, kind):
    """`kind` states moving sideways at 0.3-1 m/s, so that the friction bound of every contact is active. Rows sitting
    on the bound make the sweeps less well conditioned in float32 (with the uniform coefficient 1.0 just the same):
    the tests on these states allow SLIDING_SLACK x the usual tolerance."""
    st = random_states(N, rng, tables, kind)
    st["root"][:, 7:9] += rng.uniform(0.3, 1.0, (N, 2)) * rng.choice([-1.0, 1.0], (N, 2))
    st["root"] = st["root"].astype(np.float32).astype(np.float64)
    return st


@pytest.mark.parametrize("kind", ["stand", "mixed"])
def test_lane_program_per_env_friction(kind):
    """Per-env friction table (DR of rigid_shape_properties.friction): sole contacts (pyramid bound of the sweeps) and
    penalty contacts (viscous-Coulomb bound) against the oracle run with the same per-env coefficients; the table
    must matter (results differ from the uniform-friction run)."""
    import dataclasses
    tables = load_assets()[0]
    cfg = CoreConfig()
    N = 12
    rng = np.random.default_rng(21)
    st = sliding_states(N, rng, tables, kind)
    mu = rng.uniform(0.2, 1.3, N).astype(np.float32).astype(np.float64)
    o = PhysicsOracle(tables, dataclasses.replace(oracle_params(cfg), mu=mu))
    want = o.substep(st["root"], st["q"], st["qd"], st["tau"], st["damping"], st["armature"], st["mass_scale"])
    got = emulate_substep(tables, cfg, st, friction=mu)
    compare(st, got, want, ctx=f"{kind}: ", slack=SLIDING_SLACK)
    uniform = emulate_substep(tables, cfg, st)
    assert np.abs(uniform[3] - got[3]).max() > 1.0, "the friction table should change the contact forces"


def test_lane_program_body_wrench_and_effort_clamp():
    tables = load_assets()[0]
    cfg = CoreConfig(clamp_effort=True, gravity=(0.0, 0.0, 0.0))
    o = PhysicsOracle(tables, oracle_params(cfg))
    rng = np.random.default_rng(5)
    st = random_states(4, rng, tables, "air")
    st["tau"] = (st["tau"] * 50).astype(np.float32).astype(np.float64)
    F, T = rng.normal(0, 50, (4, 38, 3)), rng.normal(0, 5, (4, 38, 3))
    want = o.substep(st["root"], st["q"], st["qd"], st["tau"], st["damping"], st["armature"], st["mass_scale"],
                     rb_force=F, rb_torque=T)
    got = emulate_substep(tables, cfg, st, rb_force=F, rb_torque=T)
    compare(st, got, want)


def test_lane_program_short_rollout_stays_with_oracle():
    tables = load_assets()[0]
    cfg = CoreConfig()
    o = PhysicsOracle(tables, oracle_params(cfg))
    rng = np.random.default_rng(9)
    st = random_states(6, rng, tables, "stand")
    a = dict(st)
    b = dict(st)
    for _ in range(5):
        r, q, qd, c, _d = o.substep(a["root"], a["q"], a["qd"], a["tau"], a["damping"], a["armature"], a["mass_scale"])
        a.update(root=r, q=q, qd=qd)
        r2, q2, qd2, c2 = emulate_substep(tables, cfg, b)
        b.update(root=r2, q=q2, qd=qd2)
    assert np.abs(a["q"] - b["q"]).max() < 2e-5 and np.abs(a["qd"] - b["qd"]).max() < 5e-3
    assert np.abs(a["root"][:, :7] - b["root"][:, :7]).max() < 2e-5
    assert np.abs(c - c2).max() < 2e-3 * np.abs(c).max()


def test_tocabi_cta_holds_28_envs():
    """4096 envs on 148 SMs run as ONE wave only if a CTA (one per SM, 227 KiB of shared memory) holds 28 envs."""
    import ctypes as C
    from tests.physics_util import hostemu
    from isaacgymdyros_b200.core import make_model_desc
    md, keep = make_model_desc(load_assets()[0], CoreConfig())
    fn = hostemu().dyros_hostemu_cta_smem_bytes
    fn.restype = C.c_long
    need = fn(C.byref(md), 28)
    assert 0 < need <= 227 * 1024 - 1024, need  # 1 KiB stays free for static shared memory
