"""Generate the committed golden vectors by running the UNMODIFIED reference task code
(/root/reference, via oracle/ref_harness.py) on CPU torch. Run in the build container only:

    python tests/golden/make_golden.py

Each scenario records, per step: the action, the scripted "simulator outputs" injected at each
`gym.simulate` call (PhysX is absent: SURVEY fact 2), every random draw the reference made (mapped to
env-indexed rows), and the reference's outputs/state after the step. The same files drive
tests/test_oracle_pinned.py (CPU, numpy oracle) and tests/test_task_parity_gpu.py (CUDA, via the C-ABI).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_harness  # noqa: E402

STATE_KEYS = ["root_states", "dof_pos", "dof_vel", "contact_forces", "init_mocap_data_idx", "mocap_data_idx", "time",
              "qpos_noise", "qvel_noise", "qpos_pre", "target_vel", "motor_constant_scale",
              "pre_joint_velocity_states", "action_torque_pre", "contact_forces_pre", "qpos_bias", "quat_bias",
              "action_torque", "target_data_qpos", "target_data_force", "delay_idx", "simul_len", "action_log",
              "epi_len", "epi_len_log", "contact_reward_sum", "contact_reward_mean", "perturbation_count",
              "pert_duration", "pert_on", "impulse", "magnitude", "phase", "perturb_timing", "perturb_start",
              "actions", "actions_pre", "obs_history", "action_history", "total_mass", "obs_buf", "rew_buf",
              "reset_buf", "timeout_buf", "progress_buf", "randomize_buf"]


def snapshot(s):
    """Reference task attributes -> dict with oracle/task_oracle.py key names."""
    d = {}
    for k in STATE_KEYS:
        if k == "delay_idx":
            v = s.delay_idx_tensor[:, 1]
        elif k == "simul_len":
            v = s.simul_len_tensor[:, 1]
        else:
            v = getattr(s, k)
        d[k] = v.detach().clone().numpy()
    return d


class DrawRecorder:
    """Patches torch.rand / randint / normal to log every draw the reference makes."""

    def __init__(self):
        self.log = []

    def __enter__(self):
        self._orig = (torch.rand, torch.randint, torch.normal)
        rec = self

        def rand(*a, **k):
            out = rec._orig[0](*a, **k)
            rec.log.append(("rand", out.clone()))
            return out

        def randint(*a, **k):
            out = rec._orig[1](*a, **k)
            rec.log.append(("randint", out.clone()))
            return out

        def normal(*a, **k):
            out = rec._orig[2](*a, **k)
            rec.log.append(("normal", out.clone()))
            return out

        torch.rand, torch.randint, torch.normal = rand, randint, normal
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randint, torch.normal = self._orig


def run_scenario(name, N, steps, seed, *, force_perturb=False, near_timeout=False, collision_rate=0.02):
    torch.set_num_threads(1)
    s = ref_harness.make_reference_task(N, seed=seed)
    T, _, _ = ref_harness.load_reference_modules()
    rng = np.random.default_rng(seed + 1000)
    if force_perturb:
        s.perturb_start[:, 0] = True
        s.perturb_timing[:] = torch.tensor(rng.integers(1, 6, N))
    if near_timeout:
        s.progress_buf[:] = torch.tensor(rng.integers(7990, 7999, N))
        s.epi_len[:] = s.progress_buf.float()
    init = snapshot(s)
    captured = {}
    orig_start, orig_reset = s.start_perturbation, s.reset_idx
    s.start_perturbation = lambda ids: (captured.__setitem__("pert_ids", ids.clone()), orig_start(ids))[1]
    s.reset_idx = lambda ids: (captured.__setitem__("reset_ids", ids.clone()), orig_reset(ids))[1]
    sim_out = []

    def hook():
        # scripted simulator output: perturb dof state, tilt/move the root, draw contact forces
        s.dof_state.view(N, 33, 2)[..., 0] += torch.tensor(rng.normal(0, 0.01, (N, 33)), dtype=torch.float)
        s.dof_state.view(N, 33, 2)[..., 1] = torch.tensor(rng.normal(0, 0.5, (N, 33)), dtype=torch.float)
        q = s.root_states[:, 3:7] + torch.tensor(rng.normal(0, 0.03, (N, 4)), dtype=torch.float)
        s.root_states[:, 3:7] = q / q.norm(dim=-1, keepdim=True)
        s.root_states[:, 0:3] += torch.tensor(rng.normal(0, 0.002, (N, 3)), dtype=torch.float)
        s.root_states[:, 7:13] = torch.tensor(rng.normal(0, 0.3, (N, 6)), dtype=torch.float)
        cf = np.zeros((N, 38, 3), np.float32)
        for foot in (8, 16):
            on = rng.random(N) < 0.6
            cf[:, foot, 2] = on * rng.uniform(0, 1700, N)
            cf[:, foot, 0:2] = on[:, None] * rng.normal(0, 40, (N, 2))
        hit = rng.random(N) < collision_rate
        body = rng.integers(0, 38, N)
        for i in np.nonzero(hit)[0]:
            if body[i] not in (8, 16):
                cf[i, body[i]] = rng.normal(0, 30, 3)
        s.contact_forces[:] = torch.tensor(cf)
        sim_out.append({"root_states": s.root_states.clone().numpy(), "dof_pos": s.dof_pos.clone().numpy(),
                        "dof_vel": s.dof_vel.clone().numpy(), "contact_forces": s.contact_forces.clone().numpy()})

    s.gym.simulate_hook = hook
    out = {"meta_N": N, "meta_steps": steps}
    for k, v in init.items():
        out[f"init/{k}"] = v
    for t in range(steps):
        captured.clear()
        sim_out.clear()
        actions = torch.tensor(rng.uniform(-1.2, 1.2, (N, 13)), dtype=torch.float)
        s.gym.calls.clear()
        s.gym.actuation.clear()
        s.gym.applied.clear()
        with DrawRecorder() as rec:
            ref_harness.reference_step(s, actions)
        # ---- map the draw log to env-indexed rows (SURVEY A6 order)
        log = list(rec.log)
        noise = {"qpos": np.zeros((2, N, 33), np.float32), "vel": np.zeros((N, 6), np.float32),
                 "reset_f": np.zeros((N, 32), np.float32), "reset_i": np.zeros((N, 2), np.int64),
                 "pert_i": np.zeros((N, 2), np.int64), "pert_f": np.zeros((N, 1), np.float32)}
        pos = 0
        if "pert_ids" in captured:
            ids = captured["pert_ids"].flatten().numpy()
            assert log[0][0] == "randint" and log[1][0] == "randint" and log[2][0] == "rand"
            noise["pert_i"][ids, 0] = log[0][1].flatten().numpy()
            noise["pert_i"][ids, 1] = log[1][1].flatten().numpy()
            noise["pert_f"][ids, 0] = log[2][1].flatten().numpy()
            pos = 3
        for k in range(2):
            assert log[pos][0] == "normal"
            noise["qpos"][k] = log[pos][1].numpy()
            pos += 1
        ids = np.zeros(0, np.int64)
        if "reset_ids" in captured:
            ids = captured["reset_ids"].numpy()
            kinds = [l[0] for l in log[pos:pos + 10]]
            assert kinds == ["rand"] * 8 + ["randint"] * 2, kinds
            r = [l[1].numpy() for l in log[pos:pos + 10]]
            noise["reset_f"][ids, 0:12] = r[0]
            noise["reset_f"][ids, 12:15] = r[1]
            noise["reset_f"][ids, 15:17] = 0.5  # ft_bias draw: never read (T:617); placeholder keeps 32 columns
            # r[2]=ft_bias(k,2), r[3]=m_bias(N,4): unused
            noise["reset_f"][ids, 15] = r[4][:, 0]
            noise["reset_f"][ids, 16] = r[5][:, 0]
            noise["reset_f"][ids, 17] = r[6][:, 0]
            noise["reset_f"][ids, 18:30] = r[7]
            noise["reset_i"][ids, 0] = r[8][:, 0]
            noise["reset_i"][ids, 1] = r[9][:, 0]
            pos += 10
        assert log[pos][0] == "rand" and tuple(log[pos][1].shape) == (N, 6)
        noise["vel"] = log[pos][1].numpy()
        assert pos + 1 == len(log), (pos, len(log))
        out[f"s{t}/actions"] = actions.numpy()
        for j, so in enumerate(sim_out):
            for k, v in so.items():
                out[f"s{t}/sim{j}/{k}"] = v
        for k, v in noise.items():
            out[f"s{t}/noise/{k}"] = v
        # the two tensors the reference hands to the simulator (north star: PD torques within 1e-5)
        assert len(s.gym.actuation) == 2
        for j, tau in enumerate(s.gym.actuation):
            out[f"s{t}/tau{j}"] = tau.numpy().reshape(N, 33)        # dyros_dynamic_walk.py:520
        push = np.zeros((N, 3), np.float32)
        if s.gym.applied:                                           # dyros_dynamic_walk.py:493-502
            assert len(s.gym.applied) == 1
            forces, torques, space = s.gym.applied[0]
            assert space == 0 and not torques.any()
            nz = forces.clone()
            nz[:, s.pelvis_idx] = 0
            assert not nz.any(), "only the pelvis is pushed"
            push = forces[:, s.pelvis_idx].numpy().copy()
        out[f"s{t}/push"] = push
        out[f"s{t}/env_ids"] = ids.astype(np.int64)
        out[f"s{t}/stacked_rewards"] = s.extras["stacked_rewards"].clone().numpy()
        for k, v in snapshot(s).items():
            out[f"s{t}/after/{k}"] = v
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, f"{os.path.getsize(path) / 1024:.0f} KiB")


def main():
    if not ref_harness.reference_available():
        raise SystemExit("reference not mounted; golden vectors can only be generated in the build container")
    run_scenario("walk_basic", N=24, steps=8, seed=3)
    run_scenario("walk_perturb", N=16, steps=6, seed=5, force_perturb=True)
    run_scenario("walk_timeout", N=16, steps=4, seed=7, near_timeout=True, collision_rate=0.0)


if __name__ == "__main__":
    main()
