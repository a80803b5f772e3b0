import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from isaacgymdyros_b200.core import CoreConfig, DyrosCore, stable_penalty
from tests.test_humanoid_generality import humanoid, HUMANOID_CFG, STAND_Z
t = humanoid()
N = 4096
k_pen, c_pen = stable_penalty(0.0166 / 2)
VL = float(sys.argv[1]) if len(sys.argv) > 1 else 1e3
SC = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
cfg2 = CoreConfig(**{**HUMANOID_CFG, "substeps": 1, "dt": 0.0083, "penalty_stiffness": k_pen, "penalty_damping": c_pen, "dof_vel_limit": VL})
core = DyrosCore(N, "cuda:0", cfg2, tables=t, with_task=False)
core.sim_t["root_states"][:, 2] = STAND_Z
g = torch.Generator(device="cuda:0"); g.manual_seed(0)
hist = []
for s in range(600):
    tau = ((torch.rand(N, 21, device="cuda:0", generator=g) * 2 - 1) * SC * torch.tensor(t.dof_effort, dtype=torch.float32, device="cuda:0")).reshape(-1)
    core.sim_t["dof_actuation_force"].copy_(tau)
    before = (core.sim_t["root_states"].clone(), core.sim_t["dof_state"].clone(), tau.clone())
    core.simulate()
    v = core.sim_t["root_states"][:, 7:10].norm(dim=1)
    bad = (v > 8).nonzero().flatten()
    if s % 100 == 99:
        print("step", s, "vmax", float(v.max()), "wmax", float(core.sim_t["root_states"][:, 10:13].norm(dim=1).max()), "qd max", float(core.sim_t["dof_state"][:, 1].abs().max()), "n>5", int((v > 5).sum()))
    vb = before[0][:, 7:10].norm(dim=1)
    bad = ((v - vb) > 2.5).nonzero().flatten()
    if len(bad):
        e = int(bad[0])
        print("step", s, "env", e, "v", float(v[e]), "root after", core.sim_t["root_states"][e].cpu().numpy())
        np.savez("gpurun_out/blowup.npz", root=before[0][e].cpu().numpy(), dof=before[1].view(N, 21, 2)[e].cpu().numpy(),
                 tau=before[2].view(N, 21)[e].cpu().numpy(), root_after=core.sim_t["root_states"][e].cpu().numpy(),
                 dof_after=core.sim_t["dof_state"].view(N, 21, 2)[e].cpu().numpy(), contact=core.sim_t["net_contact_force"].view(N, 16, 3)[e].cpu().numpy())
        break
else:
    print("no blow-up; vmax", float(v.max()))
