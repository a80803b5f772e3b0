"""BASELINE configs[2] timing: the stock IsaacGymEnvs Humanoid (nv_humanoid.xml, dt 1/60 in 2 sub-steps) through the gym
facade (gym.simulate = one k_simulate launch with 2 sub-steps). Passive humanoids held for a quarter of a second at a
time (they are re-posed before they collapse: DESIGN.md section 4, known limits). Prints sim-steps/s.
    python tools/bench_humanoid.py [N]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from isaacgymdyros_b200 import gymapi, gymtorch

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
gym = gymapi.acquire_gym()
sp = gymapi.SimParams()
sp.dt, sp.substeps, sp.up_axis = 0.0166, 2, gymapi.UP_AXIS_Z
sp.gravity = gymapi.Vec3(0.0, 0.0, -9.81)
sp.physx.contact_offset, sp.physx.num_position_iterations, sp.physx.num_velocity_iterations = 0.02, 4, 0
sim = gym.create_sim(0, -1, gymapi.SIM_PHYSX, sp)
pp = gymapi.PlaneParams()
pp.normal = gymapi.Vec3(0.0, 0.0, 1.0)
gym.add_ground(sim, pp)
asset = gym.load_asset(sim, "../../assets", "mjcf/nv_humanoid.xml", gymapi.AssetOptions())
for i in range(N):
    env = gym.create_env(sim, gymapi.Vec3(0, 0, 0), gymapi.Vec3(0, 0, 0), 64)
    gym.create_actor(env, asset, gymapi.Transform(gymapi.Vec3(0.0, 0.0, 1.34), gymapi.Quat(0, 0, 0, 1)), "humanoid", i, 0, 0)
assert gym.prepare_sim(sim)
root = gymtorch.wrap_tensor(gym.acquire_actor_root_state_tensor(sim))
dof = gymtorch.wrap_tensor(gym.acquire_dof_state_tensor(sim))
root0, dof0 = root.clone(), dof.clone()
ids = torch.arange(N, dtype=torch.int32, device=root.device)
tau = torch.zeros(N * 21, device=root.device)
gym.set_dof_actuation_force_tensor(sim, gymtorch.unwrap_tensor(tau))
def block(steps):
    for _ in range(steps):
        gym.simulate(sim)
for _ in range(3):
    block(15)
    root.copy_(root0); dof.copy_(dof0)
    gym.set_actor_root_state_tensor_indexed(sim, gymtorch.unwrap_tensor(root), gymtorch.unwrap_tensor(ids), N)
    gym.set_dof_state_tensor_indexed(sim, gymtorch.unwrap_tensor(dof), gymtorch.unwrap_tensor(ids), N)
torch.cuda.synchronize()
K, tot = 20, 0.0
for _ in range(K):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); block(15); e1.record()
    root.copy_(root0); dof.copy_(dof0)
    gym.set_actor_root_state_tensor_indexed(sim, gymtorch.unwrap_tensor(root), gymtorch.unwrap_tensor(ids), N)
    gym.set_dof_state_tensor_indexed(sim, gymtorch.unwrap_tensor(dof), gymtorch.unwrap_tensor(ids), N)
    torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
ms = tot / (K * 15)
assert torch.isfinite(root).all()
print(f"Humanoid (16 bodies, 21 DOF), N={N}: {ms*1000:.1f} us per gym.simulate (2 sub-steps of 8.3 ms) = {N/ms/1e3:.2f} M sim-steps/s")
