"""GPU tests of the whole env step through the public API (fused dyros_task_step / DyrosDynamicWalk.step):
agreement with the CPU oracle (task restatement + fp64 physics) on injected draws, equivalence of the fused and
staged call sequences, CUDA-graph replay, and size-independent properties at the full 4096-env size."""
import numpy as np
import pytest
import torch

from isaacgymdyros_b200.core import CoreConfig
from oracle import task_oracle as O
from oracle.env_oracle import EnvOracle
from tests.golden_util import load_assets
from tests.physics_util import oracle_params
from tests.test_task_parity_gpu import inject_noise, load_state, make_core, read_state

pytestmark = pytest.mark.gpu


def test_fused_step_matches_cpu_oracle_rollout():
    """8 policy steps of 48 envs from the reset pose: physics + self-collision + PD + noise + reward + reset + obs against
    the oracle (about a minute of CPU: the oracle is dense fp64 numpy). Integer outputs exact; float tolerances account for
    fp32 vs fp64 physics over the rollout."""
    N = 48
    tables, mocap, obs_norm = load_assets()
    cfg = CoreConfig()
    rng = np.random.default_rng(11)
    ref = EnvOracle(N, tables, mocap, obs_norm, phys_params=oracle_params(cfg), rng=rng)
    ref.s["progress_buf"][:4] = 7997  # a few time-outs so that reset_idx runs inside the fused kernel
    ref.s["epi_len"][:4] = 7997
    core = make_core(N)
    load_state(core, ref.s)
    for t in range(8):
        noise = O.draw_noise(N, 2, rng)
        actions = rng.uniform(-1, 1, (N, 13)).astype(np.float32)
        want_ids = ref.step(actions, noise)
        inject_noise(core, noise)
        core.step(torch.tensor(actions, device=core.device))
        core.compact_resets()  # the fused step does not need the id list of T:554; it is produced on demand
        torch.cuda.synchronize()
        got = read_state(core)
        n = int(core.task_t["reset_count"].item())
        assert np.array_equal(core.task_t["reset_env_ids"][:n].cpu().numpy(), want_ids), f"step {t}"
        for k in ("reset_buf", "timeout_buf", "progress_buf", "mocap_data_idx", "delay_idx", "simul_len"):
            assert np.array_equal(got[k].reshape(-1), np.asarray(ref.s[k]).reshape(-1)), f"step {t} {k}"
        # fp32 kernel vs fp64 oracle over a 16-sub-step rollout with contact: positions 2e-5, velocities 2e-3; the reward
        # (14 exponential terms of those states) reaches 1.2e-3 by the 8th step (5e-4 through the 7th)
        got["root_pose"], got["root_vel"] = got["root_states"][:, :7], got["root_states"][:, 7:]
        ref.s["root_pose"], ref.s["root_vel"] = ref.s["root_states"][:, :7], ref.s["root_states"][:, 7:]
        for k, tol in (("dof_pos", 2e-5), ("dof_vel", 2e-3), ("root_pose", 2e-5), ("root_vel", 2e-3), ("rew_buf", 2e-3),
                       ("obs_buf", 1e-2)):
            err = np.abs(got[k].reshape(ref.s[k].shape) - ref.s[k]).max()
            assert err < tol, f"step {t}: {k} differs by {err}"
    core.close()


@pytest.mark.parametrize("optional_tables", [False, True])
def test_fused_step_equals_staged_sequence_bitwise(optional_tables):
    """dyros_task_step must give the same bits as the staged calls it fuses (same kernels, same order); also with the
    optional per-env friction and PD-gain tables in use."""
    N = 257
    rng = np.random.default_rng(3)
    kw = dict(dr_friction_range=(0.5, 1.2), dr_pd_gain_range=(0.8, 1.2)) if optional_tables else {}
    a, b = make_core(N, **kw), make_core(N, **kw)
    if optional_tables:
        mu = torch.tensor(rng.uniform(0.5, 1.2, N).astype(np.float32), device=a.device)
        gs = torch.tensor(rng.uniform(0.8, 1.2, (N, 2)).astype(np.float32), device=a.device)
        for c in (a, b):
            c.sim_t["contact_friction"].copy_(mu)
            c.task_t["pd_gain_scale"].copy_(gs)
    tables, mocap, obs_norm = load_assets()
    s, _c = O.new_state(N, mocap, obs_norm, np.full(N, np.float32(tables.total_mass())), tables.dof_lower,
                        tables.dof_upper, O.Params(), rng=rng)
    s["progress_buf"][::5] = 7998
    load_state(a, s)
    load_state(b, s)
    for t in range(4):
        noise = O.draw_noise(N, 2, rng)
        actions = torch.tensor(rng.uniform(-1, 1, (N, 13)).astype(np.float32), device=a.device)
        inject_noise(a, noise)
        inject_noise(b, noise)
        a.step(actions)
        b.prologue(actions)
        for k in range(2):
            b.substep_torque()
            if k == 0:
                b.sim_t["rb_force"].zero_()
                b.sim_t["rb_force"].view(N, 38, 3)[:, 0, :] = b.task_t["push_force"]
                b.simulate(apply_wrench=True)
            else:
                b.simulate()
            b.sensor_noise(k)
        b.epilogue(); b.check_termination(); b.compute_reward(); b.compact_resets(); b.reset_idx(None)
        b.compute_observations(); b.late_update(); b.end_step()
        a.compact_resets()
        torch.cuda.synchronize()
        ga, gb = read_state(a), read_state(b)
        # of contact_forces_pre the fused step keeps only the two foot rows, the ones the reward reads (T:858-859, T:904-907)
        for g in (ga, gb):
            g["contact_forces_pre"] = g["contact_forces_pre"].reshape(N, 38, 3)[:, [8, 16]]
        for k in ga:
            assert np.array_equal(ga[k], gb[k], equal_nan=True), f"step {t}: {k} differs between fused and staged"
    a.close(); b.close()


def test_prologue_physics_launch_equals_prologue_then_physics_bitwise():
    """dyros_task_prologue_physics (the first launch of the fused step: the prologue runs on the I/O warps of the physics
    kernel, restructured per slab) against dyros_task_prologue + dyros_task_physics, at a ragged size (N % 28 != 0)."""
    N = 999
    rng = np.random.default_rng(8)
    a, b = make_core(N), make_core(N)
    tables, mocap, obs_norm = load_assets()
    s, _c = O.new_state(N, mocap, obs_norm, np.full(N, np.float32(tables.total_mass())), tables.dof_lower,
                        tables.dof_upper, O.Params(), rng=rng)
    s["time"][:] = rng.uniform(0, 30, N).astype(np.float32).reshape(s["time"].shape)  # spread over the mocap cycle
    s["perturb_start"] = np.ones_like(s["perturb_start"])                            # push schedule active (T:492)
    s["epi_len"][:] = rng.integers(0, 400, N).astype(np.float32).reshape(s["epi_len"].shape)
    load_state(a, s)
    load_state(b, s)
    for t in range(3):
        noise = O.draw_noise(N, 2, rng)
        actions = torch.tensor(rng.uniform(-1.5, 1.5, (N, 13)).astype(np.float32), device=a.device)
        inject_noise(a, noise)
        inject_noise(b, noise)
        a.prologue_physics(actions)
        b.prologue(actions)
        b.task_physics()
        torch.cuda.synchronize()
        ga, gb = read_state(a), read_state(b)
        for k in ga:
            assert np.array_equal(ga[k], gb[k], equal_nan=True), f"step {t}: {k} differs"
        for x in (a, b):  # finish the step the same way on both
            x.epilogue(); x.check_termination(); x.compute_reward(); x.compact_resets(); x.reset_idx(None)
            x.compute_observations(); x.late_update(); x.end_step()
    a.close(); b.close()


def make_core(N, **kw):  # noqa: F811  (rb force tensors needed by the staged push path)
    from isaacgymdyros_b200.core import DyrosCore
    return DyrosCore(N, "cuda:0", CoreConfig(with_rb_force_tensors=True, **kw))


def test_vectask_api_rollout_properties_full_size():
    """DyrosDynamicWalk.step at N=4096 with production RNG and CUDA-graph replay: shapes/dtypes of the VecTask API,
    finite outputs, reset bookkeeping (progress_buf == 0 exactly where reset_buf == 1, compacted ids == nonzero),
    reward bounds, and that random actions make robots fall and get reset."""
    from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
    N = 4096
    env = DyrosDynamicWalk(default_cfg(N), "cuda:0")
    assert env.num_obs == 487 and env.num_acts == 13 and env.num_envs == N
    obs0 = env.reset()["obs"]
    assert obs0.shape == (N, 487) and not obs0.any()
    g = torch.Generator(device="cuda:0"); g.manual_seed(42)
    total_resets = 0
    for t in range(300):
        obs, rew, rst, extras = env.step(torch.rand(N, 13, device="cuda:0", generator=g) * 2 - 1)
        if t % 50 == 49 or t == 0:
            torch.cuda.synchronize()
            assert obs["obs"].shape == (N, 487) and obs["obs"].dtype == torch.float32
            assert rew.shape == (N,) and rst.dtype == torch.int64 and extras["time_outs"].dtype == torch.int64
            assert extras["stacked_rewards"].shape == (N, 15) and len(extras["reward_names"]) == 15
            assert torch.isfinite(obs["obs"]).all() and torch.isfinite(rew).all()
            assert torch.isfinite(env.root_states).all() and torch.isfinite(env.dof_state).all()
            assert ((env.progress_buf == 0) == (rst == 1)).all()
            env.core.compact_resets()  # id list of T:554, on demand
            n = int(env.core.task_t["reset_count"].item())
            assert torch.equal(env.core.task_t["reset_env_ids"][:n], rst.nonzero().flatten())
            assert rew.max().item() <= 2.5 and rew.min().item() >= -0.3
            assert (env.dof_vel.abs() <= 4.03 + 1e-4).all()
        total_resets += int(rst.sum().item()) if t % 10 == 9 else 0
    assert total_resets > 0, "random actions should make some robots fall within 1.2 s"
    env.close()


def test_step_reads_pinned_host_actions_like_device_actions():
    """A pinned host tensor handed to step() is read by the first kernel directly (no staging copy, no graph): same
    results, bit for bit, as the same actions on the device (CUDA-graph replay path)."""
    from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
    N = 300
    a, b = DyrosDynamicWalk(default_cfg(N), "cuda:0"), DyrosDynamicWalk(default_cfg(N), "cuda:0")
    g = torch.Generator(device="cpu"); g.manual_seed(5)
    for t in range(6):
        act = (torch.rand(N, 13, generator=g) * 2 - 1).pin_memory()
        oa, ra, sa, _ = a.step(act.to("cuda:0"))
        ob, rb, sb, _ = b.step(act)
        torch.cuda.synchronize()
        assert torch.equal(oa["obs"], ob["obs"]) and torch.equal(ra, rb) and torch.equal(sa, sb), f"step {t}"
    assert torch.equal(a.root_states, b.root_states) and torch.equal(a.dof_state, b.dof_state)
    a.close(); b.close()


def test_step_reads_a_reused_device_tensor_in_place():
    """A device tensor handed to step() again and again (the policy's output buffer) is read in place by a CUDA graph
    captured for its address from the second sighting on: same bits as the staging-copy path without graphs."""
    from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
    N = 130
    a = DyrosDynamicWalk(default_cfg(N), "cuda:0", use_cuda_graph=False)
    b = DyrosDynamicWalk(default_cfg(N), "cuda:0")
    g = torch.Generator(device="cuda:0"); g.manual_seed(8)
    bufs = [torch.zeros(N, 13, device="cuda:0") for _ in range(2)]
    for t in range(9):
        act = torch.rand(N, 13, device="cuda:0", generator=g) * 2 - 1
        buf = bufs[t % 2]
        buf.copy_(act)
        oa, ra, sa, _ = a.step(act)
        ob, rb, sb, _ = b.step(buf)
        torch.cuda.synchronize()
        assert torch.equal(oa["obs"], ob["obs"]) and torch.equal(ra, rb) and torch.equal(sa, sb), f"step {t}"
    assert len(b._inplace_graphs) == 2
    assert torch.equal(a.root_states, b.root_states) and torch.equal(a.dof_state, b.dof_state)
    ob, _, _, _ = b.step(bufs[0][:, :13].t().contiguous().t())  # non-contiguous: staging path, no new graph
    assert len(b._inplace_graphs) == 2
    a.close(); b.close()


def test_step_async_pipeline_returns_the_results_of_step_in_host_memory():
    """step_async / step_wait (up to three steps in flight, one D2H block per step on a copy stream) must hand back,
    bit for bit and in order, what the synchronous step() of a twin env returns; tickets out of the window raise."""
    from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
    N = 301  # odd: N*487 is not a multiple of 4 (scalar tail of the pack kernel)
    a, b = DyrosDynamicWalk(default_cfg(N), "cuda:0"), DyrosDynamicWalk(default_cfg(N), "cuda:0")
    for e in (a, b):  # episode ends inside the rollout: reset and (progress already at the limit, VT:325) time_outs non-zero
        e.progress_buf[:4] = 7999
        e.progress_buf[4:7] = 7996
    g = torch.Generator(device="cpu"); g.manual_seed(6)
    acts = [(torch.rand(N, 13, generator=g) * 2 - 1).pin_memory() for _ in range(7)]
    want = []
    for act in acts:
        o, r, s, ex = a.step(act)
        torch.cuda.synchronize()
        want.append((o["obs"].cpu().clone(), r.cpu().clone(), s.cpu().clone(), ex["time_outs"].cpu().clone()))
    assert any(w[3].any() for w in want) and any(w[2].any() for w in want)
    tickets = []
    for t, act in enumerate(acts):
        tickets.append(b.step_async(act if t % 2 == 0 else act.to("cuda:0")))  # pinned-host and device actions alike
        if t >= 1:
            o, r, s, ex = b.step_wait(tickets[t - 1])
            assert not o["obs"].is_cuda and o["obs"].is_pinned()
            w = want[t - 1]
            assert torch.equal(o["obs"], w[0]) and torch.equal(r, w[1]) and torch.equal(s, w[2]), f"step {t - 1}"
            assert torch.equal(ex["time_outs"], w[3]), f"step {t - 1}"
    o, r, s, ex = b.step_wait(tickets[-1])
    assert torch.equal(o["obs"], want[-1][0]) and torch.equal(s, want[-1][2])
    with pytest.raises(RuntimeError):
        b.step_wait(tickets[0])  # its slot has been reused
    with pytest.raises(RuntimeError):
        b.step_wait(len(acts))   # not submitted yet
    a.close(); b.close()


def test_large_ragged_shard_runs_multi_wave():
    """20,001 envs (715 CTAs of 28 envs: five waves on 148 SMs, last CTA ragged; state beyond the L2): finite outputs and
    exact reset bookkeeping, and env 0 evolves exactly as in a 4096-env shard stepped with the same actions (envs are
    independent: results must not depend on the shard size or the launch geometry)."""
    from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
    N, M = 20001, 4096
    big, small = DyrosDynamicWalk(default_cfg(N), "cuda:0"), DyrosDynamicWalk(default_cfg(M), "cuda:0")
    # same initial state and per-env constants for the first M envs
    for k, v in small.core.task_t.items():
        if v.dim() > 0 and v.shape[0] == M and big.core.task_t[k].shape[0] == N:
            big.core.task_t[k][:M].copy_(v)
    for k, v in small.core.sim_t.items():
        rows = v.shape[0] // M
        if v.shape[0] == rows * M and big.core.sim_t[k].shape[0] == rows * N:
            big.core.sim_t[k][:rows * M].copy_(v)
    g = torch.Generator(device="cuda:0"); g.manual_seed(13)
    for t in range(12):
        act = torch.rand(N, 13, device="cuda:0", generator=g) * 2 - 1
        ob, rb, sb, _ = big.step(act)
        os_, rs, ss, _ = small.step(act[:M].contiguous())
    torch.cuda.synchronize()
    assert torch.isfinite(ob["obs"]).all() and torch.isfinite(rb).all() and torch.isfinite(big.root_states).all()
    big.core.compact_resets()  # id list of T:554, on demand
    n = int(big.core.task_t["reset_count"].item())
    assert torch.equal(big.core.task_t["reset_env_ids"][:n], sb.nonzero().flatten())
    assert ((big.progress_buf == 0) == (sb == 1)).all()
    # envs that never reset use no shard-dependent random draws besides their own Philox streams (keyed by env index)
    assert torch.equal(big.root_states[:M], small.root_states) and torch.equal(ob["obs"][:M], os_["obs"])
    big.close(); small.close()


def test_optional_friction_and_pd_gain_randomisation():
    """BASELINE configs[3] names friction and PD-gain randomisation (commented out in the reference's yaml, CFG:89-96):
    the optional per-env tables are drawn at start, re-drawn for exactly the envs that reset, stay in range, and the
    rollout with them differs from the one without."""
    from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
    N = 512
    env = DyrosDynamicWalk(default_cfg(N, friction_range=(0.7, 1.3), pd_gain_range=(0.9, 1.1)), "cuda:0", use_cuda_graph=False)
    ref = DyrosDynamicWalk(default_cfg(N), "cuda:0", use_cuda_graph=False)
    assert "contact_friction" not in ref.core.sim_t and "pd_gain_scale" not in ref.core.task_t
    mu0, g0 = env.core.sim_t["contact_friction"].clone(), env.core.task_t["pd_gain_scale"].clone()
    assert mu0.min() >= 0.7 - 1e-6 and mu0.max() <= 1.3 + 1e-6 and mu0.std() > 0.1
    assert g0.min() >= 0.9 - 1e-6 and g0.max() <= 1.1 + 1e-6 and g0.std() > 0.03
    ids = torch.arange(0, N, 2, device="cuda:0")
    env.randomize_buf.fill_(1)  # VT:540-544: only envs whose randomize_buf reached the frequency (1) are re-drawn
    env.reset_idx(ids)
    torch.cuda.synchronize()
    mu1, g1 = env.core.sim_t["contact_friction"], env.core.task_t["pd_gain_scale"]
    assert (mu1[1::2] == mu0[1::2]).all() and (mu1[0::2] != mu0[0::2]).float().mean() > 0.99
    assert (g1[1::2] == g0[1::2]).all() and (g1[0::2] != g0[0::2]).float().mean() > 0.99
    assert mu1.min() >= 0.7 - 1e-6 and mu1.max() <= 1.3 + 1e-6 and g1.min() >= 0.9 - 1e-6 and g1.max() <= 1.1 + 1e-6
    g = torch.Generator(device="cuda:0"); g.manual_seed(3)
    for _ in range(20):
        act = torch.rand(N, 13, device="cuda:0", generator=g) * 2 - 1
        env.step(act); ref.step(act)
    torch.cuda.synchronize()
    assert torch.isfinite(env.obs_buf).all() and torch.isfinite(env.root_states).all()
    assert not torch.equal(env.dof_state, ref.dof_state)
    env.close(); ref.close()


def test_domain_randomisation_redraw_on_reset_ranges():
    from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
    from isaacgymdyros_b200.core import ARMATURE
    N = 512
    env = DyrosDynamicWalk(default_cfg(N, randomize=True), "cuda:0", use_cuda_graph=False)
    d0 = env.core.sim_t["dof_damping"].clone()
    ids = torch.arange(0, N, 2, device="cuda:0")
    env.randomize_buf.fill_(1)  # VT:540-544
    env.reset_idx(ids)
    torch.cuda.synchronize()
    d1, a1 = env.core.sim_t["dof_damping"], env.core.sim_t["dof_armature"]
    assert (d1[1::2] == d0[1::2]).all() and (d1[0::2] != d0[0::2]).float().mean() > 0.99
    assert d1.min() >= 0.1 - 1e-6 and d1.max() <= 3.0 + 1e-5
    ratio = a1 / torch.tensor(ARMATURE, device="cuda:0")
    assert ratio.min() >= 0.8 - 1e-5 and ratio.max() <= 1.2 + 1e-5
    ms = env.core.sim_t["body_mass_scale"]
    assert ms.min() >= 0.8 and ms.max() <= 1.2 and ms.std() > 0.05
    assert torch.allclose(env.total_mass, (ms * torch.tensor(env.core.tables.body_inertia[:, 0], dtype=torch.float32, device="cuda:0")).sum(1), rtol=1e-5)
    env.close()
