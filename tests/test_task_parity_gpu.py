"""GPU parity of the task kernels (through the C ABI) against (1) the committed golden vectors produced by the
unmodified reference code and (2) the numpy oracle on larger seeded inputs.

Masks, counters and compacted reset indices must be bit-exact; floats within 1e-5 relative (north star).
The simulator output is injected between the staged calls exactly where the reference calls gym.simulate.
"""
import numpy as np
import pytest
import torch

from tests.golden_util import COMPARE, SCENARIOS, Golden, assert_field, load_assets

pytestmark = pytest.mark.gpu


def make_core(N, **cfg_kw):
    from isaacgymdyros_b200.core import CoreConfig, DyrosCore
    return DyrosCore(N, "cuda:0", CoreConfig(**cfg_kw))


def load_state(core, st):
    """Oracle/golden-named state dict -> device buffers."""
    dev = core.device
    N = core.N
    T = lambda a, dt=None: torch.tensor(np.ascontiguousarray(a), device=dev) if dt is None else \
        torch.tensor(np.ascontiguousarray(a), device=dev).to(dt)
    core.sim_t["root_states"].copy_(T(st["root_states"]))
    ds = core.sim_t["dof_state"].view(N, 33, 2)
    ds[:, :, 0] = T(st["dof_pos"])
    ds[:, :, 1] = T(st["dof_vel"])
    core.sim_t["net_contact_force"].view(N, 38, 3).copy_(T(st["contact_forces"]))
    tb = core.task_t
    for k in ["time", "total_mass", "init_mocap_data_idx", "mocap_data_idx"]:
        tb[k].copy_(T(st[k]).reshape(N).to(tb[k].dtype))
    for k in ["delay_idx", "simul_len", "perturbation_count", "pert_duration", "pert_on", "impulse", "perturb_timing",
              "reset_buf", "timeout_buf", "progress_buf", "randomize_buf"]:
        tb[k].copy_(T(st[k]).to(tb[k].dtype))
    tb["perturb_start"].fill_(int(np.asarray(st["perturb_start"]).any()))
    for k in ["qpos_noise", "qvel_noise", "qpos_pre", "target_vel", "motor_constant_scale",
              "pre_joint_velocity_states", "action_torque_pre", "contact_forces_pre", "qpos_bias", "quat_bias",
              "action_torque", "target_data_qpos", "target_data_force", "action_log", "epi_len", "epi_len_log",
              "contact_reward_sum", "contact_reward_mean", "magnitude", "phase", "actions", "actions_pre", "obs_buf",
              "rew_buf"]:
        tb[k].copy_(T(st[k]).reshape(tb[k].shape))
    tb["obs_history"].copy_(T(st["obs_history"]).reshape(N, 20, 37))
    tb["action_history"].copy_(T(st["action_history"]).reshape(N, 20, 13))
    tb["obs_hist_head"].fill_(19)
    tb["act_hist_head"].fill_(19)
    if "env_origins" in st:
        tb["env_origins"].copy_(T(st["env_origins"]))


def read_state(core):
    N = core.N
    tb, s = core.task_t, core.sim_t
    out = {k: tb[k].cpu().numpy() for k in tb if k not in ("obs_history", "action_history", "mocap_data")}
    ds = s["dof_state"].view(N, 33, 2)
    out["root_states"] = s["root_states"].cpu().numpy()
    out["dof_pos"], out["dof_vel"] = ds[:, :, 0].cpu().numpy(), ds[:, :, 1].cpu().numpy()
    out["obs_history"] = core.obs_history_linear().cpu().numpy()
    out["action_history"] = core.action_history_linear().cpu().numpy()
    out["perturb_start"] = np.full((N, 1), int(tb["perturb_start"].item()))
    return out


def inject_noise(core, noise):
    dev = core.device
    core.set_noise_injection(
        qpos_normal=torch.tensor(noise["qpos"], device=dev).contiguous(),
        vel_u=torch.tensor(noise["vel"], device=dev).contiguous(),
        reset_f=torch.tensor(noise["reset_f"], device=dev).contiguous(),
        reset_i=torch.tensor(noise["reset_i"], device=dev).contiguous(),
        pert_i=torch.tensor(noise["pert_i"], device=dev).contiguous(),
        pert_f=torch.tensor(noise["pert_f"], device=dev).contiguous())


def staged_step(core, actions, sim_outputs, sim_inputs=None):
    """VT:293-344 + T:449-563 as the staged C-ABI calls; sim_outputs[k] replaces gym.simulate #k (T:525).
    `sim_inputs` (dict) receives what the kernels hand to the simulator: the actuation forces of each substep (T:520)
    and the pelvis push (T:498-502)."""
    dev, N = core.device, core.N
    core.prologue(torch.tensor(actions, device=dev).contiguous())
    if sim_inputs is not None:
        sim_inputs["push"] = core.task_t["push_force"].cpu().numpy().copy()
        sim_inputs["tau"] = []
    for k in range(2):
        core.substep_torque()
        if sim_inputs is not None:
            sim_inputs["tau"].append(core.sim_t["dof_actuation_force"].view(N, 33).cpu().numpy().copy())
        o = sim_outputs[k]
        core.sim_t["root_states"].copy_(torch.tensor(o["root_states"], device=dev))
        ds = core.sim_t["dof_state"].view(N, 33, 2)
        ds[:, :, 0] = torch.tensor(o["dof_pos"], device=dev)
        ds[:, :, 1] = torch.tensor(o["dof_vel"], device=dev)
        core.sim_t["net_contact_force"].view(N, 38, 3).copy_(torch.tensor(o["contact_forces"], device=dev))
        core.sensor_noise(k)
    core.epilogue()
    core.check_termination()
    core.compute_reward()
    core.compact_resets()
    core.reset_idx(None)
    core.compute_observations()
    core.late_update()
    core.end_step()
    torch.cuda.synchronize()
    n = int(core.task_t["reset_count"].item())
    ids = core.task_t["reset_env_ids"][:n].cpu().numpy()
    ids32 = core.task_t["reset_env_ids32"][:n].cpu().numpy()
    assert np.array_equal(ids, ids32.astype(np.int64)) and ids32.dtype == np.int32
    return ids


@pytest.mark.parametrize("name", SCENARIOS)
def test_cuda_task_kernels_match_reference_golden(name):
    g = Golden(name)
    core = make_core(g.N)
    load_state(core, g.init)
    for t, st in enumerate(g.step):
        inject_noise(core, st["noise"])
        seen = {}
        ids = staged_step(core, st["actions"], st["sim"], seen)
        assert np.array_equal(ids, st["env_ids"]), f"{name} step {t}: compacted reset ids differ"
        for j in range(2):  # simulator inputs vs the tensors the reference passed to gym (T:520, T:502)
            assert_field(f"tau{j}", seen["tau"][j], st["tau"][j], "float", ctx=f"{name} step {t} ")
        assert_field("push", seen["push"], st["push"], "float", ctx=f"{name} step {t} ")
        got = read_state(core)
        for k, kind in COMPARE.items():
            assert_field(k, got[k], st["after"][k], kind, ctx=f"{name} step {t} ")
        assert_field("stacked_rewards", got["stacked_rewards"], st["stacked_rewards"], "float", ctx=f"{name} step {t} ")
    core.close()


def scripted_sim(rng, s, N, collision_rate=0.02):
    """Same kind of scripted simulator output as tests/golden/make_golden.py, applied to an oracle state."""
    s["dof_pos"] = (s["dof_pos"] + rng.normal(0, 0.01, (N, 33))).astype(np.float32)
    s["dof_vel"] = rng.normal(0, 0.5, (N, 33)).astype(np.float32)
    q = s["root_states"][:, 3:7] + rng.normal(0, 0.03, (N, 4)).astype(np.float32)
    s["root_states"][:, 3:7] = q / np.linalg.norm(q, axis=-1, keepdims=True)
    s["root_states"][:, 0:3] += rng.normal(0, 0.002, (N, 3)).astype(np.float32)
    s["root_states"][:, 7:13] = rng.normal(0, 0.3, (N, 6)).astype(np.float32)
    cf = np.zeros((N, 38, 3), np.float32)
    for foot in (8, 16):
        on = rng.random(N) < 0.6
        cf[:, foot, 2] = on * rng.uniform(0, 1700, N)
        cf[:, foot, 0:2] = on[:, None] * rng.normal(0, 40, (N, 2))
    hit = np.nonzero(rng.random(N) < collision_rate)[0]
    body = rng.integers(0, 38, N)
    for i in hit:
        if body[i] not in (8, 16):
            cf[i, body[i]] = rng.normal(0, 30, 3)
    s["contact_forces"] = cf


@pytest.mark.parametrize("N,steps,perturb", [(1, 3, False), (33, 6, True), (4096, 4, True)])
def test_cuda_task_kernels_match_oracle(N, steps, perturb):
    """Seeded random rollouts at sizes the golden files do not cover (N=1, ragged last block, full size)."""
    from oracle import task_oracle as O
    tables, mocap, obs_norm = load_assets()
    rng = np.random.default_rng(1234 + N)
    total_mass = np.full(N, np.float32(tables.total_mass())) * rng.uniform(0.8, 1.2, N).astype(np.float32)
    s, c = O.new_state(N, mocap, obs_norm, total_mass, tables.dof_lower, tables.dof_upper, O.Params(), rng=rng)
    s["progress_buf"] = rng.integers(0, 8005, N)
    s["epi_len"] = s["progress_buf"].astype(np.float32)
    if perturb:
        s["perturb_start"][:] = True
        s["perturb_timing"] = rng.integers(1, 4, N)
    core = make_core(N)
    load_state(core, s)
    for t in range(steps):
        noise = O.draw_noise(N, 2, rng)
        actions = rng.uniform(-1.2, 1.2, (N, 13)).astype(np.float32)
        outs = []

        want_tau, want_push = [], []

        def simulate(st, tau, ext):
            want_tau.append(tau.copy())
            if ext is not None:
                want_push.append(ext.copy())
            scripted_sim(rng, st, N)
            outs.append({k: st[k].copy() for k in ("root_states", "dof_pos", "dof_vel", "contact_forces")})
        want_ids = O.step(s, c, actions, noise, simulate)
        inject_noise(core, noise)
        seen = {}
        ids = staged_step(core, actions, outs, seen)
        assert np.array_equal(ids, want_ids), f"step {t}: compacted reset ids differ"
        for j in range(2):
            assert_field(f"tau{j}", seen["tau"][j], want_tau[j], "float", ctx=f"N={N} step {t} ")
        assert_field("push", seen["push"], want_push[0], "float", ctx=f"N={N} step {t} ")
        got = read_state(core)
        for k, kind in COMPARE.items():
            assert_field(k, got[k], s[k], kind, ctx=f"N={N} step {t} ")
        assert_field("stacked_rewards", got["stacked_rewards"], s["stacked_rewards"], "float", ctx=f"N={N} step {t} ")
    core.close()


def test_cuda_substep_torque_with_pd_gain_scale():
    """DyrosTaskBuffers.pd_gain_scale (optional per-env scales of Kp, Kv in T:506): both torque stages (warp-per-env
    kernel, slab stage of the fused physics launch) against the oracle, float32 products in the same order: exact."""
    from oracle import task_oracle as O
    tables, mocap, obs_norm = load_assets()
    N = 61
    rng = np.random.default_rng(77)
    s, c = O.new_state(N, mocap, obs_norm, np.full(N, np.float32(tables.total_mass())), tables.dof_lower,
                       tables.dof_upper, O.Params(), rng=rng)
    s["dof_pos"] = (s["dof_pos"] + rng.normal(0, 0.1, (N, 33))).astype(np.float32)
    s["dof_vel"] = rng.normal(0, 1, (N, 33)).astype(np.float32)
    s["target_data_qpos"] = (s["dof_pos"] + rng.normal(0, 0.2, (N, 33))).astype(np.float32)
    s["action_torque"] = rng.normal(0, 50, (N, 12)).astype(np.float32)
    s["pd_gain_scale"] = rng.uniform(0.8, 1.2, (N, 2)).astype(np.float32)
    core = make_core(N, dr_pd_gain_range=(0.8, 1.2))
    load_state(core, s)
    core.task_t["pd_gain_scale"].copy_(torch.tensor(s["pd_gain_scale"], device=core.device))
    want = O.substep_torque(s, c)
    core.substep_torque()
    torch.cuda.synchronize()
    got = core.sim_t["dof_actuation_force"].view(N, 33).cpu().numpy()
    assert np.array_equal(got, want)
    unscaled = dict(s)
    del unscaled["pd_gain_scale"]
    assert np.abs(O.substep_torque(unscaled, c)[:, 12:] - want[:, 12:]).max() > 1e-3
    core.close()


def test_actions_validation_and_errors():
    from isaacgymdyros_b200 import native
    core = make_core(4)
    with pytest.raises(native.DyrosError):
        core.prologue(torch.zeros(4, 12, device="cuda:0"))
    with pytest.raises(native.DyrosError):
        core.sensor_noise(2)
    core.close()


@pytest.mark.parametrize("N", [1, 1023, 1024, 1025, 4096, 70001])
def test_reset_compaction_is_nonzero_at_tile_boundaries(N):
    """T:554 `reset_buf.nonzero()` by the multi-CTA ballot / prefix-scan kernel (tiles of 1024 envs): ascending int64
    ids, their int32 copy (T:737,745) and the count, bit-exact, for empty, full and random masks; two launches in a row
    (the kernel cleans up its own tile state)."""
    core = make_core(N)
    g = torch.Generator(device="cuda:0"); g.manual_seed(N)
    for mask in (torch.zeros(N, device="cuda:0"), torch.ones(N, device="cuda:0"),
                 (torch.rand(N, device="cuda:0", generator=g) < 0.3).float(),
                 (torch.rand(N, device="cuda:0", generator=g) < 0.01).float()):
        core.task_t["reset_buf"].copy_(mask.long() * 7)  # any non-zero value counts
        for rep in range(2):
            core.task_t["reset_env_ids"].fill_(-1)
            core.compact_resets()
            torch.cuda.synchronize()
            want = torch.nonzero(core.task_t["reset_buf"]).flatten()
            n = int(core.task_t["reset_count"].item())
            assert n == want.numel()
            assert torch.equal(core.task_t["reset_env_ids"][:n], want)
            assert torch.equal(core.task_t["reset_env_ids32"][:n].long(), want)
            assert core.task_t["reset_env_ids32"].dtype == torch.int32
    core.close()
