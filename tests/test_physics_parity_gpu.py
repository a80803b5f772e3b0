"""GPU parity of dyros_simulate (CUDA, float32, O(n) recursions, through the C ABI) against the dense fp64 oracle
on identical states. PhysX itself is absent from the reference checkout (parity with it unpinned, see
oracle/physics_oracle.py); the per-quantity tolerances below are on single-step state deltas."""
import numpy as np
import pytest
import torch

from isaacgymdyros_b200.core import CoreConfig
from oracle.physics_oracle import PhysicsOracle
from tests.golden_util import load_assets
from tests.physics_util import oracle_params, random_states
from tests.test_physics_emulation import compare

pytestmark = pytest.mark.gpu


def load_sim_state(core, st):
    dev, N = core.device, core.N
    T = lambda a: torch.tensor(np.ascontiguousarray(a, dtype=np.float32), device=dev)
    core.sim_t["root_states"].copy_(T(st["root"]))
    ds = core.sim_t["dof_state"].view(N, 33, 2)
    ds[:, :, 0] = T(st["q"])
    ds[:, :, 1] = T(st["qd"])
    core.sim_t["dof_actuation_force"].copy_(T(st["tau"]).reshape(-1))
    core.sim_t["dof_damping"].copy_(T(st["damping"]))
    core.sim_t["dof_armature"].copy_(T(st["armature"]))
    core.sim_t["body_mass_scale"].copy_(T(st["mass_scale"]))


def read_sim_state(core):
    N = core.N
    ds = core.sim_t["dof_state"].view(N, 33, 2)
    f = lambda t: t.cpu().numpy().astype(np.float64)
    return f(core.sim_t["root_states"]), f(ds[:, :, 0]), f(ds[:, :, 1]), f(core.sim_t["net_contact_force"].view(N, 38, 3))


@pytest.mark.parametrize("program", ["roles", "lanes"])
@pytest.mark.parametrize("kind,N,seed", [("air", 7, 0), ("stand", 64, 1), ("mixed", 300, 2)])
def test_cuda_simulate_matches_dense_oracle(kind, N, seed, program):
    """Both mappings of gym.simulate to the GPU (DyrosSimDesc.physics_program): one lane per env (default) and 8 lanes
    per env. (PhysicsOracle.substep is the ground-contact physics alone; the self-collision pass that follows it in
    dyros_simulate has its own oracle and tests, tests/test_self_collision.py, and is switched off here: the random
    'air' and 'mixed' poses are contorted enough to self-intersect.)"""
    from isaacgymdyros_b200.core import DyrosCore
    tables = load_assets()[0]
    cfg = CoreConfig(with_rb_force_tensors=True, physics_program=program, self_collision=False)
    o = PhysicsOracle(tables, oracle_params(cfg))
    rng = np.random.default_rng(seed)
    st = random_states(N, rng, tables, kind)
    F, Tq = rng.normal(0, 30, (N, 38, 3)), rng.normal(0, 3, (N, 38, 3))
    want = o.substep(st["root"], st["q"], st["qd"], st["tau"], st["damping"], st["armature"], st["mass_scale"],
                     rb_force=F, rb_torque=Tq)
    core = DyrosCore(N, "cuda:0", cfg)
    load_sim_state(core, st)
    core.sim_t["rb_force"].copy_(torch.tensor(F, dtype=torch.float32).reshape(-1, 3))
    core.sim_t["rb_torque"].copy_(torch.tensor(Tq, dtype=torch.float32).reshape(-1, 3))
    core.simulate(apply_wrench=True)
    torch.cuda.synchronize()
    compare(st, read_sim_state(core), want, ctx=f"{kind}: ")
    core.close()


@pytest.mark.parametrize("kind", ["stand", "mixed"])
def test_cuda_simulate_per_env_friction(kind):
    """DyrosSimBuffers.contact_friction: per-env friction of the sole contacts (sweeps) and of the penalty contacts
    against the oracle run with the same coefficients, on states that slide sideways."""
    import dataclasses
    from isaacgymdyros_b200.core import DyrosCore
    from tests.test_physics_emulation import SLIDING_SLACK, sliding_states
    tables = load_assets()[0]
    N = 96
    cfg = CoreConfig(dr_friction_range=(0.2, 1.3), self_collision=False)
    rng = np.random.default_rng(33)
    st = sliding_states(N, rng, tables, kind)
    mu = rng.uniform(0.2, 1.3, N).astype(np.float32)
    o = PhysicsOracle(tables, dataclasses.replace(oracle_params(cfg), mu=mu.astype(np.float64)))
    want = o.substep(st["root"], st["q"], st["qd"], st["tau"], st["damping"], st["armature"], st["mass_scale"])
    core = DyrosCore(N, "cuda:0", cfg)
    load_sim_state(core, st)
    core.sim_t["contact_friction"].copy_(torch.tensor(mu))
    core.simulate()
    torch.cuda.synchronize()
    got = read_sim_state(core)
    compare(st, got, want, ctx=f"{kind}: ", slack=SLIDING_SLACK)
    core.sim_t["contact_friction"].fill_(1.0)  # the table is what the kernel uses
    load_sim_state(core, st)
    core.simulate()
    torch.cuda.synchronize()
    assert np.abs(read_sim_state(core)[3] - got[3]).max() > 1.0
    core.close()


def test_cuda_standing_carries_weight_on_feet_only():
    """Physical invariant on the GPU at the full size: 4096 robots held at the reset pose by a stiff PD settle with
    the ground reaction equal to their weight, on bodies 8 and 16 only."""
    from isaacgymdyros_b200.core import DyrosCore, INIT_DOF_POS, KP, KV
    N = 4096
    core = DyrosCore(N, "cuda:0", CoreConfig())
    dev = core.device
    kp, kv = torch.tensor(KP, device=dev), torch.tensor(KV, device=dev)
    q0 = torch.tensor(INIT_DOF_POS, device=dev)
    ds = core.sim_t["dof_state"].view(N, 33, 2)
    for s in range(400):
        core.sim_t["dof_actuation_force"].copy_((kp * (q0 - ds[:, :, 0]) - kv * ds[:, :, 1]).reshape(-1))
        core.simulate()
    torch.cuda.synchronize()
    cf = core.sim_t["net_contact_force"].view(N, 38, 3)
    fz = cf[:, :, 2].sum(1)
    w = 104.48712 * 9.81
    assert torch.isfinite(core.sim_t["root_states"]).all()
    assert (fz - w).abs().max().item() < 0.05 * w
    others = [b for b in range(38) if b not in (8, 16)]
    assert cf[:, others].abs().max().item() == 0.0
    assert (core.sim_t["root_states"][:, 2] - 0.93).abs().max().item() < 0.01
    core.close()


def test_cuda_rigid_body_state_matches_oracle():
    from isaacgymdyros_b200.core import DyrosCore
    tables = load_assets()[0]
    cfg = CoreConfig(with_rigid_body_state=True)
    o = PhysicsOracle(tables, oracle_params(cfg))
    N = 50
    st = random_states(N, np.random.default_rng(4), tables, "mixed")
    core = DyrosCore(N, "cuda:0", cfg)
    load_sim_state(core, st)
    core.refresh_rigid_body_state()
    torch.cuda.synchronize()
    got = core.sim_t["rigid_body_state"].view(N, 38, 13).cpu().numpy().astype(np.float64)
    want = o.rigid_body_state(st["root"], st["q"], st["qd"])
    assert np.abs(got[..., :3] - want[..., :3]).max() < 5e-6
    sign = np.sign((got[..., 3:7] * want[..., 3:7]).sum(-1, keepdims=True))
    assert np.abs(got[..., 3:7] * sign - want[..., 3:7]).max() < 5e-6
    assert np.abs(got[..., 7:] - want[..., 7:]).max() < 2e-5
    core.close()


def test_lanes_program_fused_step_equals_staged_and_tracks_the_default_program():
    """The multi-lane variant behind the whole fused step (dyros_task_step): bitwise equal to its own staged sequence
    pieces where both exist (physics launch alone vs inside the step), and within float rounding of the default
    program over a short rollout from identical states and draws."""
    from isaacgymdyros_b200.core import DyrosCore
    N = 100
    cores = {p: DyrosCore(N, "cuda:0", CoreConfig(physics_program=p, perturb=False), seed=7) for p in ("roles", "lanes")}
    g = torch.Generator(device="cuda:0"); g.manual_seed(5)
    for t in range(6):
        act = torch.rand(N, 13, device="cuda:0", generator=g) * 2 - 1
        for c in cores.values():
            c.step(act)
    torch.cuda.synchronize()
    a, b = cores["roles"], cores["lanes"]
    same = a.task_t["reset_buf"] == b.task_t["reset_buf"]  # (a contact on the edge of a threshold may flip one env)
    assert same.float().mean().item() >= 0.97
    ds = lambda c: c.sim_t["dof_state"].view(N, 33, 2)[same]
    assert torch.allclose(ds(a), ds(b), rtol=0, atol=5e-3)
    assert torch.allclose(a.sim_t["root_states"][same], b.sim_t["root_states"][same], rtol=0, atol=2e-3)
    for c in cores.values():
        c.close()
