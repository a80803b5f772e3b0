"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference task code from /root/reference.

Only usable in the build container (the GPU box has no /root/reference). Used by
tests/golden/make_golden.py to generate the committed golden vectors and by the CPU tests that pin
oracle/task_oracle.py. Never imported by the product package.

Recipe: SURVEY.md Appendix C. The reference's `gym.simulate` (PhysX) is absent, so the fake gym's
`simulate` hook lets a test inject "what the simulator returned" per substep.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("DYROS_REFERENCE_ROOT", "/root/reference")
_PY = os.path.join(REF_ROOT, "python")
_ENVS = os.path.join(_PY, "IsaacGymEnvs", "isaacgymenvs")
_ASSETS = os.path.join(_PY, "IsaacGymEnvs", "assets")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_ENVS, "tasks", "dyros_dynamic_walk.py"))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_CACHE = {}


def load_reference_modules():
    """Returns (task_module, torch_utils, torch_jit_utils) loaded from the reference by file path."""
    if "T" in _CACHE:
        return _CACHE["T"], _CACHE["TU"], _CACHE["JU"]
    if not hasattr(np, "float"):
        np.float = float  # isaacgym/torch_utils.py:135 default argument
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.split(".")[0] in ("isaacgym", "isaacgymenvs")}
    for k in saved:
        del sys.modules[k]

    def pkg(name):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        return m

    ig = pkg("isaacgym")
    gymapi = types.ModuleType("isaacgym.gymapi")
    gymapi.ENV_SPACE = 0
    gymtorch = types.ModuleType("isaacgym.gymtorch")
    gymtorch.wrap_tensor = lambda t: t
    gymtorch.unwrap_tensor = lambda t: t
    sys.modules["isaacgym.gymapi"], sys.modules["isaacgym.gymtorch"] = gymapi, gymtorch
    ig.gymapi, ig.gymtorch = gymapi, gymtorch
    for n in ("isaacgymenvs", "isaacgymenvs.utils", "isaacgymenvs.tasks", "isaacgymenvs.tasks.base",
              "isaacgymenvs.cfg", "isaacgymenvs.cfg.terrain"):
        pkg(n)
    vt = types.ModuleType("isaacgymenvs.tasks.base.vec_task")

    class VecTask:  # empty base: the harness supplies every attribute
        pass

    vt.VecTask = VecTask
    sys.modules["isaacgymenvs.tasks.base.vec_task"] = vt
    terr = types.ModuleType("isaacgymenvs.utils.terrain")
    terr.Terrain = object
    sys.modules["isaacgymenvs.utils.terrain"] = terr
    TU = _load("isaacgym.torch_utils", os.path.join(_PY, "isaacgym", "torch_utils.py"))
    JU = _load("isaacgymenvs.utils.torch_jit_utils", os.path.join(_ENVS, "utils", "torch_jit_utils.py"))
    _load("isaacgymenvs.cfg.terrain.terrain_cfg", os.path.join(_ENVS, "cfg", "terrain", "terrain_cfg.py"))
    T = _load("isaacgymenvs.tasks.dyros_dynamic_walk", os.path.join(_ENVS, "tasks", "dyros_dynamic_walk.py"))
    # leave the stubs registered only under the reference's own names; restore whatever was there before
    _CACHE.update(T=T, TU=TU, JU=JU)
    for k in [k for k in sys.modules if k.split(".")[0] in ("isaacgym", "isaacgymenvs")]:
        del sys.modules[k]
    sys.modules.update({k: v for k, v in saved.items() if v is not None})
    return T, TU, JU


class FakeGym:
    """Every gym.* call is a logged no-op returning True; `simulate` runs an optional hook."""

    def __init__(self, masses):
        self.calls = []
        self.simulate_hook = None
        self._masses = masses
        self.actuation = []   # tensors handed to set_dof_actuation_force_tensor (dyros_dynamic_walk.py:520), in call order
        self.applied = []     # (forces, torques, space) handed to apply_rigid_body_force_tensors (:502)

    def set_dof_actuation_force_tensor(self, sim, tensor):
        self.calls.append("set_dof_actuation_force_tensor")
        self.actuation.append(tensor.detach().clone())
        return True

    def apply_rigid_body_force_tensors(self, sim, forces, torques, space):
        self.calls.append("apply_rigid_body_force_tensors")
        self.applied.append((forces.detach().clone(), torques.detach().clone(), space))
        return True

    def get_actor_rigid_body_properties(self, env, handle):
        self.calls.append("get_actor_rigid_body_properties")
        return [types.SimpleNamespace(mass=float(m)) for m in self._masses[int(env)]]

    def simulate(self, sim):
        self.calls.append("simulate")
        if self.simulate_hook is not None:
            self.simulate_hook()

    def __getattr__(self, name):
        def f(*a, **k):
            self.calls.append(name)
            return True
        return f


KP = [2000.0, 5000.0, 4000.0, 3700.0, 3200.0, 3200.0, 2000.0, 5000.0, 4000.0, 3700.0, 3200.0, 3200.0,
      6000.0, 10000.0, 10000.0, 400.0, 1000.0, 400.0, 400.0, 400.0, 400.0, 100.0, 100.0, 100.0, 100.0,
      400.0, 1000.0, 400.0, 400.0, 400.0, 400.0, 100.0, 100.0]
KV = [15.0, 50.0, 20.0, 25.0, 24.0, 24.0, 15.0, 50.0, 20.0, 25.0, 24.0, 24.0, 200.0, 100.0, 100.0,
      10.0, 28.0, 10.0, 10.0, 10.0, 10.0, 3.0, 3.0, 2.0, 2.0, 10.0, 28.0, 10.0, 10.0, 10.0, 10.0, 3.0, 3.0]


def make_reference_task(num_envs: int, seed: int = 0, body_masses=None, dof_limits=None):
    """Build a reference DyrosDynamicWalk with __init__ bypassed and every A2 attribute assigned exactly
    as tasks/dyros_dynamic_walk.py:58-195 / vec_task.py:233-256 would (CPU, float32)."""
    T, TU, JU = load_reference_modules()
    torch.manual_seed(seed)
    N, nd, nb = num_envs, 33, 38
    s = T.DyrosDynamicWalk.__new__(T.DyrosDynamicWalk)
    dev = "cpu"
    s.device = dev
    s.cfg = {"env": {"envSpacing": 5}}
    s.num_envs = N
    s.num_dof, s.num_bodies, s.num_actions, s.num_action = nd, nb, 13, 13
    s.num_obs_his, s.num_obs_skip, s.num_single_step_obs = 10, 2, 37
    s.randomize, s.randomization_params = False, {}
    s.death_cost, s.termination_height, s.initial_height = 0.0, 0.6, 0.93
    s.max_episode_length_s = 32
    s.max_episode_length = 32 / (0.002 * 2)
    s.perturb = True
    s.terrain_cfg = sys_terrain_cfg(T)
    s.custom_origins = False
    s.init_done = True
    s.extras = {}
    s.sim, s.viewer = None, None
    s.render = lambda: None
    s.envs = list(range(N))
    s.humanoid_handles = [0] * N
    if body_masses is None:
        body_masses = np.tile(np.load(os.path.join(os.path.dirname(__file__), "..", "isaacgymdyros_b200",
                                                   "assets", "tocabi_tables.npz"))["body_inertia"][:, 0], (N, 1))
    s.gym = FakeGym(body_masses)
    s.Kp = torch.tensor(KP, dtype=torch.float) / 9.0
    s.Kv = torch.tensor(KV, dtype=torch.float) / 3.0
    s.root_states = torch.zeros(N, 13)
    s.root_states[:, 2] = 0.93
    s.root_states[:, 6] = 1.0
    s.dof_state = torch.zeros(N * nd, 2)
    s.dof_pos = s.dof_state.view(N, nd, 2)[..., 0]
    s.dof_vel = s.dof_state.view(N, nd, 2)[..., 1]
    s.initial_dof_pos = torch.zeros(N, nd)
    s.initial_dof_pos[:, 0:] = torch.tensor([0.0, 0.0, -0.24, 0.6, -0.36, 0.0, 0.0, 0.0, -0.24, 0.6, -0.36, 0.0,
                                             0.0, 0.0, 0.0, 0.3, 0.3, 1.5, -1.27, -1.0, 0.0, -1.0, 0.0, 0.0, 0.0,
                                             -0.3, -0.3, -1.5, 1.27, 1.0, 0.0, 1.0, 0.0])
    s.initial_dof_vel = torch.zeros(N, nd)
    s.contact_forces = torch.zeros(N * nb, 3).view(N, nb, 3)
    s.dof_pos[:] = s.initial_dof_pos[:]
    s.init_mocap_data_idx = torch.zeros(N, 1, dtype=torch.long)
    s.mocap_data_idx = torch.zeros(N, 1, dtype=torch.long)
    mocap = np.genfromtxt(os.path.join(_ASSETS, "DeepMimic", "processed_data_tocabi_walk.txt"), encoding="ascii")
    s.mocap_data = torch.tensor(mocap, dtype=torch.float)
    s.mocap_data_num = int(s.mocap_data.shape[0] - 1)
    s.mocap_cycle_dt = 0.0005
    s.mocap_cycle_period = s.mocap_data_num * s.mocap_cycle_dt
    s.time = torch.zeros(N, 1)
    s.dt, s.skipframe = 0.002, 2
    s.dt_policy = s.dt * s.skipframe
    s.policy_freq_scale = 1 / (s.dt_policy * 250)
    s.sim_time_scale = s.dt / 0.0005
    s.qpos_noise = torch.zeros_like(s.dof_pos)
    s.qvel_noise = torch.zeros_like(s.dof_vel)
    s.qpos_pre = torch.zeros_like(s.dof_pos)
    vel_mag = torch.rand(N, 1) * 0.8
    s.target_vel = torch.cat([vel_mag, vel_mag * 0.0], dim=1)
    s.motor_constant_scale = torch.rand(N, 12) * 0.4 + 0.8
    s.obs_mean = torch.tensor(np.genfromtxt(os.path.join(_ASSETS, "Data", "obs_mean_fixed.txt"), encoding="ascii"),
                              dtype=torch.float)
    s.obs_var = torch.tensor(np.genfromtxt(os.path.join(_ASSETS, "Data", "obs_variance_fixed.txt"), encoding="ascii"),
                             dtype=torch.float)
    s.pre_joint_velocity_states = s.dof_vel.clone()
    s.action_torque_pre = torch.zeros(N, 12)
    s.contact_forces_pre = s.contact_forces.clone()
    s.qpos_bias = torch.rand(N, 12) * 6.28 / 100 - 3.14 / 100
    s.quat_bias = torch.rand(N, 3) * 6.28 / 150 - 3.14 / 150
    s.ft_bias = torch.rand(N, 2) * 200.0 - 100.0
    s.m_bias = torch.rand(N, 4) * 20.0 - 10.0
    s.action_torque = torch.zeros(N, 12)
    s.target_data_qpos = torch.zeros(N, nd)
    s.target_data_force = torch.zeros(N, 2)
    s.delay_idx_tensor = torch.zeros(N, 2, dtype=torch.long)
    s.simul_len_tensor = torch.zeros(N, 2, dtype=torch.long)
    s.delay_idx_tensor[:, 1] = 1
    s.delay_idx_tensor[:, 0] = torch.arange(N)
    s.simul_len_tensor[:, 0] = torch.arange(N)
    s.action_log = torch.zeros(N, round(0.01 / s.dt) + 1, 12)
    s.epi_len = torch.zeros(N)
    s.epi_len_log = torch.zeros(N)
    s.contact_reward_sum = torch.zeros(N)
    s.contact_reward_mean = torch.zeros(N)
    s.perturbation_count = torch.zeros(N, dtype=torch.long)
    s.pert_duration = torch.randint(low=1, high=100, size=(N, 1)).squeeze(-1)
    s.pert_on = torch.zeros(N, dtype=torch.bool)
    s.impulse = torch.zeros(N, dtype=torch.long)
    s.magnitude = torch.zeros(N)
    s.phase = torch.zeros(N)
    s.perturb_timing = torch.ones(N, dtype=torch.long)
    s.perturb_start = torch.zeros(N, 1, dtype=torch.bool)
    s.actions = torch.zeros(N, 13)
    s.actions_pre = torch.zeros(N, 13)
    s.obs_history = torch.zeros(N, 10 * 2 * 37)
    s.action_history = torch.zeros(N, 10 * 2 * 13)
    s.action_high = torch.tensor([333, 232, 263, 289, 222, 166, 333, 232, 263, 289, 222, 166, 303, 303, 303,
                                  64, 64, 64, 64, 23, 23, 10, 10, 10, 10, 64, 64, 64, 64, 23, 23, 10, 10])
    s.pelvis_idx, s.left_foot_idx, s.right_foot_idx = 0, 8, 16
    s.non_feet_idxs = [i for i in range(nb) if i not in (8, 16)]
    s.initial_root_states = torch.tensor([0.0, 0.0, 0.93, 0, 0, 0, 1.0, 0, 0, 0, 0, 0, 0])
    if dof_limits is None:
        z = np.load(os.path.join(os.path.dirname(__file__), "..", "isaacgymdyros_b200", "assets", "tocabi_tables.npz"))
        dof_limits = (z["dof_lower"], z["dof_upper"])
    s.dof_limits_lower = torch.tensor(dof_limits[0], dtype=torch.float)
    s.dof_limits_upper = torch.tensor(dof_limits[1], dtype=torch.float)
    s._get_env_origins()
    s.total_mass = torch.tensor(body_masses, dtype=torch.float).sum(dim=1, keepdim=True)
    # VecTask buffers (vec_task.py:233-256)
    s.obs_buf = torch.zeros(N, 487)
    s.rew_buf = torch.zeros(N)
    s.reset_buf = torch.ones(N, dtype=torch.long)
    s.timeout_buf = torch.zeros(N, dtype=torch.long)
    s.progress_buf = torch.zeros(N, dtype=torch.long)
    s.randomize_buf = torch.zeros(N, dtype=torch.long)
    return s


def sys_terrain_cfg(T):
    return T.TerrainCfg()


def reference_step(s, actions: torch.Tensor):
    """VecTask.step body for this task (vec_task.py:293-344) around the reference's own pre/post methods."""
    a = torch.clamp(actions, -1.0, 1.0)
    s.pre_physics_step(a)
    s.timeout_buf = torch.where(s.progress_buf >= s.max_episode_length - 1, torch.ones_like(s.timeout_buf),
                                torch.zeros_like(s.timeout_buf))
    s.post_physics_step()
    return s.obs_buf, s.rew_buf, s.reset_buf
