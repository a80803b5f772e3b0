"""Helpers shared by the CPU (oracle) and GPU (CUDA) golden-vector tests."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENARIOS = ["walk_basic", "walk_perturb", "walk_timeout"]
ASSETS = os.path.join(os.path.dirname(GOLDEN_DIR), "..", "isaacgymdyros_b200", "assets")

# after-step fields compared: name -> "exact" (integer / mask / index work) or "float"
COMPARE = {
    "reset_buf": "exact", "timeout_buf": "exact", "progress_buf": "exact", "randomize_buf": "exact",
    "mocap_data_idx": "exact", "init_mocap_data_idx": "exact", "delay_idx": "exact", "simul_len": "exact",
    "perturbation_count": "exact", "pert_duration": "exact", "pert_on": "exact", "impulse": "exact",
    "perturb_timing": "exact", "perturb_start": "exact",
    "obs_buf": "float", "rew_buf": "float", "root_states": "float", "dof_pos": "float", "dof_vel": "float",
    "time": "float", "qpos_noise": "float", "qvel_noise": "float", "qpos_pre": "float", "target_vel": "float",
    "motor_constant_scale": "float", "pre_joint_velocity_states": "float", "action_torque_pre": "float",
    "contact_forces_pre": "float", "qpos_bias": "float", "quat_bias": "float", "action_torque": "float",
    "target_data_qpos": "float", "target_data_force": "float", "action_log": "float", "epi_len": "float",
    "epi_len_log": "float", "contact_reward_sum": "float", "contact_reward_mean": "float", "magnitude": "float",
    "phase": "float", "actions": "float", "actions_pre": "float", "obs_history": "float", "action_history": "float",
}
RTOL, ATOL = 1e-5, 1e-6  # north star: 1e-5 relative (fp32); atol covers values that are analytically ~0


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.N, self.steps = int(z["meta_N"]), int(z["meta_steps"])
        self.init = {k[5:]: z[k] for k in z.files if k.startswith("init/")}
        self.step = []
        for t in range(self.steps):
            p = f"s{t}/"
            d = {"actions": z[p + "actions"], "env_ids": z[p + "env_ids"], "stacked_rewards": z[p + "stacked_rewards"],
                 "sim": [{k.split("/")[-1]: z[k] for k in z.files if k.startswith(f"{p}sim{j}/")} for j in range(2)],
                 # what the reference handed to set_dof_actuation_force_tensor (T:520, per substep) and the pelvis row
                 # of apply_rigid_body_force_tensors (T:498-502)
                 "tau": [z[f"{p}tau{j}"] for j in range(2)], "push": z[p + "push"],
                 "noise": {k.split("/")[-1]: z[k] for k in z.files if k.startswith(p + "noise/")},
                 "after": {k.split("/")[-1]: z[k] for k in z.files if k.startswith(p + "after/")}}
            self.step.append(d)


def load_assets():
    from isaacgymdyros_b200.model.tables import ModelTables
    t = ModelTables.load(os.path.join(ASSETS, "tocabi_tables.npz"))
    return t, np.load(os.path.join(ASSETS, "mocap_walk.npy")), np.load(os.path.join(ASSETS, "obs_norm.npy"))


def assert_field(name, got, want, kind, ctx=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape or got.size == want.size, f"{ctx}{name}: shape {got.shape} vs {want.shape}"
    got = got.reshape(want.shape)
    if kind == "exact":
        assert np.array_equal(got.astype(np.int64), want.astype(np.int64)), f"{ctx}{name}: integer/mask mismatch"
    else:
        both_nan = np.isnan(got) & np.isnan(want)
        ok = np.isclose(got, want, rtol=RTOL, atol=ATOL) | both_nan
        if not ok.all():
            i = np.argwhere(~ok)[0]
            raise AssertionError(f"{ctx}{name}: {int((~ok).sum())} mismatches, first at {tuple(i)}: "
                                 f"got {got[tuple(i)]!r} want {want[tuple(i)]!r}")
