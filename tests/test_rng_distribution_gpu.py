"""Distribution tests of the PRODUCTION Philox streams: the parity tests inject every draw, so the streams
themselves are checked here at N = 4096: range, mean, variance and independence of each draw site, and the integer
ranges of the randint sites exactly as the reference's Python forms them (T:441, T:652, T:665)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
N = 4096


def corr(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    a, b = a - a.mean(), b - b.mean()
    return float((a * b).mean() / (a.std() * b.std() + 1e-300))


def test_sensor_noise_stream_is_clamped_gaussian_and_independent():
    """T:528: clamp(normal(0, 0.00016/3), +-0.00016) per DOF, sub-step and step; 200 steps x 2 sub-steps x 33 DOF."""
    from isaacgymdyros_b200.core import CoreConfig, DyrosCore
    core = DyrosCore(N, "cuda:0", CoreConfig(perturb=False))
    sig = 0.00016 / 3.0
    ds = core.sim_t["dof_state"].view(N, 33, 2)
    samples = []
    for step in range(100):
        for k in range(2):
            core.sensor_noise(k)
            samples.append((core.task_t["qpos_noise"] - ds[:, :, 0]).clone())
        core.end_step()  # bumps the Philox epoch
    x = torch.stack(samples)  # (200, N, 33); exact in float32 up to the rounding of pos + n at |pos| <= 1.5
    assert float(x.abs().max()) <= 0.00016 + 2e-7
    assert abs(float(x.mean())) < 3 * sig / np.sqrt(x.numel()) + 1e-9
    # variance of a normal clamped at 3 sigma: 0.99501 sigma^2
    assert abs(float(x.double().std()) / sig - np.sqrt(0.99501)) < 0.01
    at_bound = float((x.abs() > 0.00016 - 2e-7).float().mean())
    assert 0.0015 < at_bound < 0.0045  # 2 * (1 - Phi(3)) = 0.0027
    # independence: the two DOFs of a Box-Muller pair, sub-steps, consecutive steps, neighbouring envs
    assert abs(corr(x[:, :, 0], x[:, :, 1])) < 0.01
    assert abs(corr(x[0::2], x[1::2])) < 0.005
    assert abs(corr(x[0:-2], x[2:])) < 0.005
    assert abs(corr(x[:, :-1], x[:, 1:])) < 0.005
    core.close()


def test_velocity_noise_stream_is_uniform():
    """T:766: rand(N,6) * 0.05 - 0.025 added to the root velocities of the newest observation frame."""
    from isaacgymdyros_b200.core import CoreConfig, DyrosCore
    core = DyrosCore(N, "cuda:0", CoreConfig(perturb=False))
    mean, var = core.task_t["obs_mean"][31:37], core.task_t["obs_var"][31:37]
    us = []
    for step in range(60):
        core.sim_t["root_states"][:, 7:13].normal_(0, 0.3)
        core.compute_observations()
        nv = core.task_t["obs_buf"][:, 37 * 9 + 31:37 * 10]  # newest frame = history position 19 (T:789-791)
        v = nv * torch.sqrt(var + 1e-8) + mean
        us.append(((v - core.sim_t["root_states"][:, 7:13]) + 0.025) / 0.05)
        core.end_step()
    u = torch.stack(us).double()
    assert float(u.min()) > -1e-3 and float(u.max()) < 1 + 1e-3
    assert abs(float(u.mean()) - 0.5) < 0.002 and abs(float(u.var()) - 1 / 12) < 0.001
    assert abs(corr(u[:, :, 0], u[:, :, 1])) < 0.005 and abs(corr(u[:-1], u[1:])) < 0.005
    core.close()


def test_reset_draws_ranges_and_fresh_draws_per_explicit_reset():
    """T:615-665 through dyros_task_reset_idx, 40 explicit resets of all envs inside ONE step epoch (every call must
    draw afresh: reset_seq is part of the Philox counter)."""
    from isaacgymdyros_b200.core import CoreConfig, DyrosCore
    core = DyrosCore(N, "cuda:0", CoreConfig(perturb=True, randomize=True))
    ids = torch.arange(N, device="cuda:0", dtype=torch.int64)
    t, s = core.task_t, core.sim_t
    rec = {k: [] for k in ("delay_idx", "perturb_timing", "qpos_bias", "quat_bias", "motor_constant_scale", "target_vel",
                           "init_mocap_data_idx", "dof_damping", "dof_armature")}
    for r in range(40):
        t["randomize_buf"].fill_(1)
        core.reset_idx(ids)
        for k in rec:
            rec[k].append((s[k] if k in s else t[k]).clone())
    torch.cuda.synchronize()
    R = {k: torch.stack(v) for k, v in rec.items()}
    d = R["delay_idx"]
    assert int(d.min()) == 2 and int(d.max()) == 5          # randint(1+int(0.002/dt), 1+round(0.01/dt)) = [2, 6)
    freq = torch.bincount(d.flatten().long(), minlength=6)[2:6].double() / d.numel()
    assert float((freq - 0.25).abs().max()) < 0.01
    pt = R["perturb_timing"]
    assert int(pt.min()) >= 0 and int(pt.max()) == 1999     # randint(0, int(8/dt_policy)) = [0, 2000)
    assert abs(float(pt.double().mean()) - 999.5) < 6
    qb = R["qpos_bias"].double()
    assert float(qb.min()) >= -0.0314 - 1e-7 and float(qb.max()) <= 0.0314 + 1e-7 and abs(float(qb.mean())) < 1e-4
    assert abs(float(qb.std()) - 0.0628 / np.sqrt(12)) < 2e-4
    quat = R["quat_bias"].double()
    assert float(quat.abs().max()) <= 3.14 / 150 + 1e-7
    mc = R["motor_constant_scale"].double()
    assert float(mc.min()) >= 0.8 and float(mc.max()) <= 1.2 + 1e-6 and abs(float(mc.mean()) - 1.0) < 1e-3
    tv = R["target_vel"].double()
    assert float(tv[..., 0].min()) >= 0 and float(tv[..., 0].max()) <= 0.8 and float(tv[..., 1].abs().max()) == 0.0
    im = R["init_mocap_data_idx"]
    assert set(im.unique().tolist()) == {0, 1800} and abs(float((im == 0).double().mean()) - 0.5) < 0.01
    dd = R["dof_damping"].double()
    assert float(dd.min()) >= 0.1 and float(dd.max()) <= 3.0 + 1e-6 and abs(float(dd.mean()) - 1.55) < 0.01
    # fresh draws on every explicit reset of the same env within one epoch (they used to be identical)
    assert float((R["qpos_bias"][0] == R["qpos_bias"][1]).float().mean()) < 0.01
    assert float((R["dof_damping"][3] == R["dof_damping"][4]).float().mean()) < 0.01
    assert abs(corr(R["qpos_bias"][:-1], R["qpos_bias"][1:])) < 0.005
    assert abs(corr(R["qpos_bias"][..., 0], R["motor_constant_scale"][..., 0])) < 0.005
    core.close()


def test_push_draws_ranges():
    """T:438-443: impulse randint(50,250), duration randint(int(0.1/dt_policy), int(1/dt_policy)) = [25, 250),
    phase U[0, 2 pi), magnitude = impulse / (duration * dt_policy)."""
    from isaacgymdyros_b200.core import CoreConfig, DyrosCore
    core = DyrosCore(N, "cuda:0", CoreConfig(perturb=True))
    t = core.task_t
    t["perturb_start"].fill_(1)
    actions = torch.zeros(N, 13, device="cuda:0")
    imp, dur, ph, mag = [], [], [], []
    for r in range(30):
        t["epi_len"].fill_(0.0)
        t["perturb_timing"].fill_(0)     # epi_len % 2000 == perturb_timing: every env starts a push (T:495)
        t["pert_on"].fill_(0)
        t["perturbation_count"].fill_(0)
        core.prologue(actions)
        imp.append(t["impulse"].clone()); dur.append(t["pert_duration"].clone())
        ph.append(t["phase"].clone()); mag.append(t["magnitude"].clone())
        core.end_step()
    torch.cuda.synchronize()
    imp, dur, ph, mag = torch.stack(imp), torch.stack(dur), torch.stack(ph).double(), torch.stack(mag)
    assert int(imp.min()) == 50 and int(imp.max()) == 249
    assert int(dur.min()) == 25 and int(dur.max()) == 249
    assert float(ph.min()) >= 0 and float(ph.max()) < 2 * np.pi and abs(float(ph.mean()) - np.pi) < 0.02
    want = imp.float() / (dur.float() * np.float32(0.004))
    assert torch.allclose(mag, want, rtol=1e-6)
    push = t["push_force"]
    assert torch.allclose(push[:, 0], mag[-1] * torch.cos(ph[-1].float()), rtol=1e-5, atol=1e-3)
    assert abs(corr(imp.double(), dur.double())) < 0.01
    core.close()


def test_unsupported_dt_is_refused():
    """The actuation-delay ring is compiled for round(0.01/dt)+1 == 6 (T:166): another dt must fail loudly."""
    from isaacgymdyros_b200 import native
    from isaacgymdyros_b200.core import CoreConfig, DyrosCore
    with pytest.raises(native.DyrosError, match="unsupported"):
        DyrosCore(8, "cuda:0", CoreConfig(dt=0.001))
