"""The gym-level boundary (gymapi / gymtorch facade) driven the way the reference task drives Isaac Gym
(tasks/dyros_dynamic_walk.py:199-385 setup, :504-530 physics loop, :720-748 indexed resets)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def build_sim(N):
    from isaacgymdyros_b200 import gymapi, gymtorch
    gym = gymapi.acquire_gym()
    sp = gymapi.SimParams()
    sp.dt, sp.substeps, sp.up_axis = 0.002, 1, gymapi.UP_AXIS_Z
    sp.gravity = gymapi.Vec3(0.0, 0.0, -9.81)
    sp.use_gpu_pipeline = True
    sp.physx.contact_offset, sp.physx.max_depenetration_velocity = 0.002, 10.0
    sim = gym.create_sim(0, -1, gymapi.SIM_PHYSX, sp)
    assert sim is not None
    pp = gymapi.PlaneParams()
    pp.normal = gymapi.Vec3(0.0, 0.0, 1.0)
    gym.add_ground(sim, pp)
    ao = gymapi.AssetOptions()
    ao.angular_damping, ao.max_angular_velocity = 0.0, 100.0
    asset = gym.load_asset(sim, "../assets", "mjcf/dyros_tocabi/xml/dyros_tocabi.xml", ao)
    assert gym.get_asset_rigid_body_count(asset) == 38 and gym.get_asset_dof_count(asset) == 33
    assert gym.find_asset_rigid_body_index(asset, "L_Foot_Link") == 8
    assert gym.find_asset_rigid_body_index(asset, "R_Foot_Link") == 16
    envs = []
    for i in range(N):
        env = gym.create_env(sim, gymapi.Vec3(0, 0, 0), gymapi.Vec3(0, 0, 0), 4)
        pose = gymapi.Transform(gymapi.Vec3(5.0 * i, 0.0, 0.93), gymapi.Quat(0, 0, 0, 1))
        h = gym.create_actor(env, asset, pose, "humanoid", i, 0, 0)
        props = gym.get_actor_dof_properties(env, h)
        props["damping"] = 0.1
        props["velocity"] = 4.03
        props["armature"] = 0.5
        assert gym.set_actor_dof_properties(env, h, props)
        envs.append(env)
    assert gym.prepare_sim(sim)
    return gym, gymapi, gymtorch, sim, envs


def test_gym_tensor_api_loop_matches_core_and_indexed_reset():
    from isaacgymdyros_b200.core import INIT_DOF_POS
    N = 6
    gym, gymapi, gymtorch, sim, envs = build_sim(N)
    root = gymtorch.wrap_tensor(gym.acquire_actor_root_state_tensor(sim))
    dof = gymtorch.wrap_tensor(gym.acquire_dof_state_tensor(sim))
    contact = gymtorch.wrap_tensor(gym.acquire_net_contact_force_tensor(sim))
    rb = gymtorch.wrap_tensor(gym.acquire_rigid_body_state_tensor(sim))
    assert root.shape == (N, 13) and dof.shape == (N * 33, 2) and contact.shape == (N * 38, 3) and rb.shape == (N * 38, 13)
    assert gymtorch.wrap_tensor(gym.acquire_dof_state_tensor(sim)).data_ptr() == dof.data_ptr()
    assert torch.allclose(root[:, 0].cpu(), torch.arange(N) * 5.0) and (root[:, 2] == 0.93).all()
    masses = [p.mass for p in gym.get_actor_rigid_body_properties(envs[0], 0)]
    assert abs(sum(masses) - 104.48712) < 1e-4
    dof.view(N, 33, 2)[:, :, 0] = torch.tensor(INIT_DOF_POS, device=dof.device)
    assert gym.set_dof_state_tensor(sim, gymtorch.unwrap_tensor(dof))
    q0 = dof.view(N, 33, 2)[:, :, 0].clone()
    kp = torch.full((33,), 2000.0, device=dof.device)
    for _ in range(150):  # T:504-526 pattern
        tau = (kp * (q0 - dof.view(N, 33, 2)[:, :, 0]) - 30.0 * dof.view(N, 33, 2)[:, :, 1]).reshape(-1).contiguous()
        assert gym.set_dof_actuation_force_tensor(sim, gymtorch.unwrap_tensor(tau))
        gym.simulate(sim)
        assert gym.refresh_dof_state_tensor(sim)
    gym.fetch_results(sim, True)
    assert gym.refresh_net_contact_force_tensor(sim) and gym.refresh_rigid_body_state_tensor(sim)
    assert gym.get_frame_count(sim) == 150
    fz = contact.view(N, 38, 3)[:, :, 2].sum(1)
    assert ((fz - 104.48712 * 9.81).abs() < 0.08 * 104.48712 * 9.81).all()
    assert torch.allclose(rb.view(N, 38, 13)[:, 0, :3], root[:, :3], atol=1e-6)  # body 0 is the root
    # a push through apply_rigid_body_force_tensors acts for one step on the pelvis
    f = torch.zeros(N * 38, 3, device=dof.device)
    f.view(N, 38, 3)[:, 0, 0] = 2000.0
    v_before = root[:, 7].clone()
    assert gym.apply_rigid_body_force_tensors(sim, gymtorch.unwrap_tensor(f), None, gymapi.ENV_SPACE)
    gym.simulate(sim)
    dv1 = (root[:, 7] - v_before).clone()
    v_before = root[:, 7].clone()
    gym.simulate(sim)
    dv2 = root[:, 7] - v_before
    assert (dv1 > 0.01).all() and (dv2.abs() < 0.5 * dv1).all()
    # indexed reset: full tensors + int32 actor ids (T:737-746)
    ids = torch.tensor([1, 4], dtype=torch.int32, device=dof.device)
    root[ids.long(), 2] = 1.5
    dof.view(N, 33, 2)[ids.long()] = 0.0
    assert gym.set_actor_root_state_tensor_indexed(sim, gymtorch.unwrap_tensor(root), gymtorch.unwrap_tensor(ids), 2)
    assert gym.set_dof_state_tensor_indexed(sim, gymtorch.unwrap_tensor(dof), gymtorch.unwrap_tensor(ids), 2)
    assert not gym.set_dof_state_tensor_indexed(sim, gymtorch.unwrap_tensor(dof), gymtorch.unwrap_tensor(ids.long()), 2)
    gym.simulate(sim)
    assert (root[ids.long(), 2] > 1.4).all() and (contact.view(N, 38, 3)[ids.long()] == 0).all()
    with pytest.raises(Exception):
        gymtorch.unwrap_tensor(dof.view(N, 33, 2)[:, :, 0])  # non-contiguous (gymtorch.py:98-99)
    gym.destroy_sim(sim)


def test_collision_filter_of_create_actor_switches_self_collision():
    """create_actor(..., group, filter, seg): filter 0 = the actor's shapes collide with each other (T:354), a non-zero mask
    shared by all its shapes switches that off (assets.rst.txt:107-109). Seen at the API: two launches per simulate vs one."""
    from isaacgymdyros_b200 import gymapi
    out = {}
    for flt in (0, 1):
        gym = gymapi.acquire_gym()
        sp = gymapi.SimParams()
        sp.dt, sp.substeps, sp.up_axis, sp.gravity = 0.002, 1, gymapi.UP_AXIS_Z, gymapi.Vec3(0.0, 0.0, -9.81)
        sp.use_gpu_pipeline = True
        sim = gym.create_sim(0, -1, gymapi.SIM_PHYSX, sp)
        pp = gymapi.PlaneParams()
        pp.normal = gymapi.Vec3(0.0, 0.0, 1.0)
        gym.add_ground(sim, pp)
        asset = gym.load_asset(sim, "../assets", "mjcf/dyros_tocabi/xml/dyros_tocabi.xml", gymapi.AssetOptions())
        for i in range(4):
            env = gym.create_env(sim, gymapi.Vec3(0, 0, 0), gymapi.Vec3(0, 0, 0), 2)
            gym.create_actor(env, asset, gymapi.Transform(gymapi.Vec3(5.0 * i, 0.0, 0.93), gymapi.Quat(0, 0, 0, 1)), "humanoid", i, flt, 0)
        assert gym.prepare_sim(sim)
        out[flt] = sim.core.cfg.self_collision
        sim.core.close()
    assert out == {0: True, 1: False}


def test_create_sim_failure_returns_none():
    from isaacgymdyros_b200 import gymapi
    gym = gymapi.acquire_gym()
    assert gym.create_sim(0, -1, gymapi.SIM_FLEX, gymapi.SimParams()) is None  # caller quits (vec_task.py:270-273)


def test_humanoid_through_the_gym_api():
    """BASELINE configs[2] at the gym level: load_asset / create_actor / prepare_sim / tensor API / simulate on the stock
    Humanoid MJCF (Humanoid.yaml: dt 1/60 in 2 sub-steps, contact_offset 0.02)."""
    from isaacgymdyros_b200 import gymapi, gymtorch
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
    assert gym.get_asset_rigid_body_count(asset) == 16 and gym.get_asset_dof_count(asset) == 21
    # force sensors at the feet and the actuator list, as tasks/humanoid.py:159-168 sets them up
    feet_idx = [gym.find_asset_rigid_body_index(asset, "right_foot"), gym.find_asset_rigid_body_index(asset, "left_foot")]
    for b in feet_idx:
        gym.create_asset_force_sensor(asset, b, gymapi.Transform())
    assert len(gym.get_asset_actuator_properties(asset)) == 21
    N = 64
    for i in range(N):
        env = gym.create_env(sim, gymapi.Vec3(0, 0, 0), gymapi.Vec3(0, 0, 0), 8)
        h = gym.create_actor(env, asset, gymapi.Transform(gymapi.Vec3(2.0 * i, 0.0, 1.34), gymapi.Quat(0, 0, 0, 1)), "humanoid", i, 0, 0)
        gym.enable_actor_dof_force_sensors(env, h)
    assert gym.prepare_sim(sim)
    sensors = gymtorch.wrap_tensor(gym.acquire_force_sensor_tensor(sim))
    dof_force = gymtorch.wrap_tensor(gym.acquire_dof_force_tensor(sim))
    assert sensors.shape == (N * 2, 6) and dof_force.shape == (N * 21,)
    root = gymtorch.wrap_tensor(gym.acquire_actor_root_state_tensor(sim))
    dof = gymtorch.wrap_tensor(gym.acquire_dof_state_tensor(sim))
    contact = gymtorch.wrap_tensor(gym.acquire_net_contact_force_tensor(sim))
    assert root.shape == (N, 13) and dof.shape == (N * 21, 2) and contact.shape == (N * 16, 3)
    tau = torch.zeros(N * 21, device=root.device)
    for _ in range(15):  # a quarter of a second: the humanoids drop 5 cm onto their feet (passive, they collapse later)
        gym.set_dof_actuation_force_tensor(sim, gymtorch.unwrap_tensor(tau))
        gym.simulate(sim)
    gym.fetch_results(sim, True)
    assert torch.isfinite(root).all() and torch.isfinite(dof).all()
    feet = [gym.find_asset_rigid_body_index(asset, "right_foot"), gym.find_asset_rigid_body_index(asset, "left_foot")]
    cf = contact.view(N, 16, 3)
    assert (cf[:, feet, 2].sum(1) > 50.0).all()  # standing on the soles
    assert (root[:, 2] > 1.0).all() and (root[:, 2] < 1.4).all()
    # refresh_dof_force_tensor / refresh_force_sensor_tensor (tasks/humanoid.py:243-245) against their definitions
    tau = torch.randn(N * 21, device=root.device)
    gym.set_dof_actuation_force_tensor(sim, gymtorch.unwrap_tensor(tau))
    assert gym.refresh_dof_force_tensor(sim) and gym.refresh_force_sensor_tensor(sim) and gym.refresh_rigid_body_state_tensor(sim)
    torch.cuda.synchronize()
    t = sim.asset.tables
    damp = sim.core.sim_t["dof_damping"].reshape(-1)
    stiff = torch.tensor(t.dof_stiffness, dtype=torch.float32, device=root.device).repeat(N)
    assert torch.allclose(dof_force, tau - damp * dof[:, 1] - stiff * dof[:, 0], rtol=1e-5, atol=1e-5)
    rb = gymtorch.wrap_tensor(gym.acquire_rigid_body_state_tensor(sim)).view(N, 16, 13)
    q = rb[:, feet_idx, 3:7]                                          # xyzw of the two feet
    F = cf[:, feet_idx, :]
    qv, qw = q[..., :3], q[..., 3:4]                                  # rotate F by the inverse of q
    tq = 2.0 * torch.cross(-qv, F, dim=-1)
    want = F + qw * tq + torch.cross(-qv, tq, dim=-1)
    got = sensors.view(N, 2, 6)
    assert torch.allclose(got[..., :3], want, rtol=1e-4, atol=1e-3) and float(got[..., 3:].abs().max()) == 0.0
    assert float(got[..., 2].min()) > 10.0                            # the soles are flat on the ground: local z carries the weight
    gym.destroy_sim(sim)
