"""The drop-in claim at gym level: the reference's UNMODIFIED task code --
tasks/dyros_dynamic_walk.py::DyrosDynamicWalk (T:22) on tasks/base/vec_task.py::VecTask (VT:155), with the reference's
own isaacgym/torch_utils.py, gymutil.py, terrain_utils.py -- constructed and stepped on this repo's `isaacgym.gymapi` /
`gymtorch` / `gymdeps` (isaacgymdyros_b200/compat), with domain randomisation on, so that VecTask.apply_randomizations
(VT:519-733) drives the facade's property getters / setters through gymutil.apply_random_samples.

Runs where /root/reference exists (the build container: no GPU), so the native core behind the facade is replaced by a
CPU stand-in whose gym.simulate is the fp64 oracle (tests/fake_core.py): what is tested is the Python surface of the
facade -- every gym.* call the reference makes, shapes, dtypes, aliasing of the acquired tensors, int32 indexed setters,
property arrays -- not CUDA numerics (those are the -m gpu tests). Environment shims only: numpy >= 2 removed np.float /
np.Inf (the reference targets numpy 1.x) and OpenAI `gym` is not installed (VecTask needs gym.spaces.Box only)."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch
import yaml

REF = os.environ.get("DYROS_REFERENCE_ROOT", "/root/reference")
ENVS = os.path.join(REF, "python", "IsaacGymEnvs", "isaacgymenvs")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(ENVS, "tasks", "dyros_dynamic_walk.py")),
                                reason="reference checkout not mounted")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture
def reference_task_class(monkeypatch):
    for attr, val in (("float", float), ("Inf", np.inf)):  # numpy 1.x names the reference uses (TU:135, VT:90)
        if not hasattr(np, attr):
            monkeypatch.setattr(np, attr, val, raising=False)
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("isaacgym", "isaacgymenvs", "gym")}
    for k in saved:
        del sys.modules[k]
    monkeypatch.setenv("DYROS_ISAACGYM_PY", os.path.join(REF, "python", "isaacgym"))
    monkeypatch.syspath_prepend(os.path.join(ROOT, "isaacgymdyros_b200", "compat"))
    from isaacgymdyros_b200.tasks.spaces import Box
    gym_stub, spaces = types.ModuleType("gym"), types.ModuleType("gym.spaces")
    spaces.Box = Box
    gym_stub.spaces, gym_stub.Space = spaces, object
    sys.modules["gym"], sys.modules["gym.spaces"] = gym_stub, spaces
    import isaacgym  # this repo's compat package; torch_utils / gymutil / terrain_utils resolve to the reference's files
    from isaacgym import gymapi, gymtorch, gymutil, torch_utils  # noqa: F401
    assert gymapi.__file__.startswith(ROOT) and gymtorch.__file__.startswith(ROOT)
    assert torch_utils.__file__.startswith(REF) and gymutil.__file__.startswith(REF)
    for n in ("isaacgymenvs", "isaacgymenvs.utils", "isaacgymenvs.tasks", "isaacgymenvs.tasks.base", "isaacgymenvs.cfg",
              "isaacgymenvs.cfg.terrain"):
        m = types.ModuleType(n)
        m.__path__ = []
        sys.modules[n] = m
    _load("isaacgymenvs.utils.torch_jit_utils", os.path.join(ENVS, "utils", "torch_jit_utils.py"))
    _load("isaacgymenvs.cfg.terrain.terrain_cfg", os.path.join(ENVS, "cfg", "terrain", "terrain_cfg.py"))
    _load("isaacgymenvs.utils.terrain", os.path.join(ENVS, "utils", "terrain.py"))
    _load("isaacgymenvs.tasks.base.vec_task", os.path.join(ENVS, "tasks", "base", "vec_task.py"))
    T = _load("isaacgymenvs.tasks.dyros_dynamic_walk", os.path.join(ENVS, "tasks", "dyros_dynamic_walk.py"))
    # the native core behind the facade -> CPU stand-in (tests/fake_core.py)
    from isaacgymdyros_b200 import gymapi as facade
    from tests.fake_core import OracleCore
    monkeypatch.setattr(facade, "_core_factory", OracleCore)
    monkeypatch.setattr(facade, "_device_available", lambda: True)
    monkeypatch.setattr(facade, "_GYM", None)
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "python", "IsaacGymEnvs", "isaacgymenvs"))  # T:112-142 read assets relative to the cwd
    try:
        yield T.DyrosDynamicWalk
    finally:
        os.chdir(cwd)
        for k in [k for k in sys.modules if k.split(".")[0] in ("isaacgym", "isaacgymenvs", "gym")]:
            del sys.modules[k]
        sys.modules.update(saved)


def reference_cfg(num_envs):
    cfg = yaml.safe_load(open(os.path.join(ENVS, "cfg", "task", "DyrosDynamicWalk.yaml")))
    cfg["env"]["numEnvs"] = num_envs
    cfg["env"]["envSpacing"] = 5
    cfg["physics_engine"] = "physx"                       # cfg/config.yaml:20
    cfg["sim"]["use_gpu_pipeline"] = False                # sim_device=cpu pipeline=cpu (BASELINE configs[0])
    cfg["sim"]["physx"]["use_gpu"] = False
    cfg["sim"]["physx"]["num_threads"] = 4
    cfg["rl_device"] = "cpu"
    return {k: v for k, v in cfg.items()}


def test_unmodified_reference_task_runs_on_the_facade(reference_task_class):
    N = 4
    np.random.seed(0)
    torch.manual_seed(0)
    cfg = reference_cfg(N)
    assert cfg["task"]["randomize"] is True
    env = reference_task_class(cfg, "cpu", -1, True)
    from isaacgymdyros_b200 import gymapi as facade
    core = env.sim.core
    # ---- what __init__ built through the facade (T:24-195, VT:157-203)
    assert env.num_envs == N and env.num_dof == 33 and env.num_bodies == 38 and env.num_obs == 487 and env.num_acts == 13
    assert (env.pelvis_idx, env.left_foot_idx, env.right_foot_idx) == (0, 8, 16)        # T:304-306
    assert env.root_states.data_ptr() == core.sim_t["root_states"].data_ptr()            # acquire_* alias the sim's buffers
    assert env.dof_state.data_ptr() == core.sim_t["dof_state"].data_ptr()
    assert env.contact_forces.shape == (N, 38, 3)
    assert torch.allclose(env.dof_pos, env.initial_dof_pos)                              # set_dof_state_tensor T:104-107
    assert env.reset_buf.dtype == torch.long and bool((env.reset_buf == 1).all())       # VT:248
    # first_randomization (T:218-219 -> VT:536-538): every env got mass x U[0.8,1.2] (setup only), damping 0.1+U[0,2.9],
    # armature x U[0.8,1.2] through get/set_actor_*_properties and gymutil.apply_random_samples
    d, a = core.sim_t["dof_damping"], core.sim_t["dof_armature"]
    from isaacgymdyros_b200.core import ARMATURE
    assert float(d.min()) >= 0.1 and float(d.max()) <= 3.0 + 1e-6 and float(d.std()) > 0.3
    ratio = a / torch.tensor(ARMATURE)
    assert float(ratio.min()) >= 0.8 - 1e-6 and float(ratio.max()) <= 1.2 + 1e-6 and float(ratio.std()) > 0.03
    ms = core.sim_t["body_mass_scale"]
    assert float(ms.min()) >= 0.8 - 1e-6 and float(ms.max()) <= 1.2 + 1e-6 and float(ms.std()) > 0.03
    base = torch.tensor(core.tables.body_inertia[:, 0], dtype=torch.float32)
    assert torch.allclose(env.total_mass[:, 0], (ms * base).sum(1), rtol=1e-5)           # T:221-225
    # ---- VecTask.step x 3 with random actions (VT:293-344 -> T:449-563): 2 simulate calls per step
    for t in range(3):
        actions = torch.rand(N, 13) * 2 - 1
        obs, rew, reset, extras = env.step(actions)
        assert obs["obs"].shape == (N, 487) and torch.isfinite(obs["obs"]).all()
        assert rew.shape == (N,) and torch.isfinite(rew).all() and reset.dtype == torch.long
        assert extras["stacked_rewards"].shape == (N, 15) and len(extras["reward_names"]) == 15
    assert core.simulate_calls == 6
    assert float(core.sim_t["net_contact_force"].view(N, 38, 3)[:, [8, 16], 2].sum()) > 100.0  # the robots stand on their soles
    # ---- a reset through reset_idx (T:598-669): DR re-drawn for the reset envs only, from the ORIGINAL properties
    d0 = core.sim_t["dof_damping"].clone()
    env.progress_buf[1] = 7999                                                          # time-out on the next step
    obs, rew, reset, extras = env.step(torch.zeros(N, 13))
    assert int(reset[1]) == 1 and int(extras["time_outs"][1]) == 1
    d1 = core.sim_t["dof_damping"]
    assert not torch.equal(d1[1], d0[1]) and float(d1[1].min()) >= 0.1 and float(d1[1].max()) <= 3.0 + 1e-6
    still = [i for i in range(N) if int(reset[i]) == 0]
    assert still and all(torch.equal(d1[i], d0[i]) for i in still)
    assert int(env.progress_buf[1]) == 0 and float(env.root_states[1, 2]) == pytest.approx(0.93)


def test_unmodified_reference_humanoid_task_runs_on_the_facade(reference_task_class):
    """BASELINE configs[2] (generality): the stock IsaacGymEnvs Humanoid task (tasks/humanoid.py:41), unmodified, on the
    facade: nv_humanoid.xml through the MJCF importer, get_asset_actuator_properties, create_asset_force_sensor,
    acquire_force_sensor_tensor / acquire_dof_force_tensor and their refreshes (humanoid.py:80-86,159-168,196,243-245).
    (In this fork VecTask.step never calls gym.simulate for stock tasks -- VT:313-319 is commented out, T:1-6 -- so the
    test steps the simulator itself between two task steps.)"""
    H = _load("isaacgymenvs.tasks.humanoid", os.path.join(ENVS, "tasks", "humanoid.py"))
    cfg = yaml.safe_load(open(os.path.join(ENVS, "cfg", "task", "Humanoid.yaml")))
    N = 3
    cfg["env"]["numEnvs"] = N
    cfg["physics_engine"] = "physx"
    cfg["sim"]["use_gpu_pipeline"] = False
    cfg["sim"]["physx"]["use_gpu"] = False
    cfg["sim"]["physx"]["num_threads"] = 4
    cfg["sim"]["physx"]["num_subscenes"] = 0
    cfg["sim"]["physx"]["max_gpu_contact_pairs"] = 1024
    cfg["task"]["randomize"] = False
    cfg["rl_device"] = "cpu"
    for k, v in list(cfg["sim"].items()):  # the yaml carries hydra interpolations for values config.yaml supplies
        if isinstance(v, str) and v.startswith("${"):
            cfg["sim"][k] = {"use_gpu_pipeline": False}.get(k, 0)
    cfg["sim"]["dt"], cfg["sim"]["substeps"] = 0.0166, 2
    env = H.Humanoid(cfg, "cpu", -1, True)
    core = env.sim.core
    assert env.num_dof == 21 and env.num_bodies == 16 and env.num_obs == 108 and env.num_acts == 21
    assert env.vec_sensor_tensor.shape == (N, 12) and env.dof_force_tensor.shape == (N, 21)
    assert len(env.motor_efforts) == 21 and float(env.max_motor_effort) > 0
    obs, rew, reset, extras = env.step(torch.zeros(N, 21))          # stock task step: no simulate in this fork
    assert obs["obs"].shape == (N, 108) and torch.isfinite(obs["obs"]).all()
    for _ in range(3):                                               # ... so drive the simulator, then step the task again
        env.gym.simulate(env.sim)
    obs, rew, reset, extras = env.step(torch.zeros(N, 21))
    assert core.simulate_calls == 3 and torch.isfinite(obs["obs"]).all() and torch.isfinite(rew).all()
    assert torch.isfinite(env.vec_sensor_tensor).all() and torch.isfinite(env.dof_force_tensor).all()
