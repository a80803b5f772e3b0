"""TEST INFRASTRUCTURE ONLY. A CPU stand-in for isaacgymdyros_b200.core.DyrosCore behind the gym facade: the same
`sim_t` tensor dictionary (CPU torch tensors) and the methods gymapi.Gym calls, with gym.simulate done by the dense
fp64 oracle (oracle/physics_oracle.py). It exists so that the PYTHON surface of the facade (every gym.* call, tensor
shape, dtype and ownership rule the reference task relies on) can be exercised in the build container, which has the
reference but no GPU. Numerics of the CUDA path are covered by the -m gpu tests, not here."""
import numpy as np
import torch

from oracle.physics_oracle import PhysicsOracle
from tests.physics_util import oracle_params


class OracleCore:
    def __init__(self, num_envs, device="cpu", cfg=None, tables=None, seed=42, rank=0, with_task=True):
        self.N, self.cfg, self.tables = int(num_envs), cfg, tables
        self.device = torch.device("cpu")
        self.with_task = False
        nd, nb, N = tables.num_dofs, tables.num_bodies, self.N
        self.nd, self.nb = nd, nb
        z = lambda *s: torch.zeros(*s, dtype=torch.float32)
        self.sim_t = {"root_states": z(N, 13), "dof_state": z(N * nd, 2), "net_contact_force": z(N * nb, 3),
                      "dof_actuation_force": z(N * nd), "dof_damping": z(N, nd), "dof_armature": z(N, nd),
                      "body_mass_scale": torch.ones(N, nb), "rigid_body_state": z(N * nb, 13),
                      "rb_force": z(N * nb, 3), "rb_torque": z(N * nb, 3)}
        self.sim_t["root_states"][:, 6] = 1.0
        self.task_t = {}
        self.oracle = PhysicsOracle(tables, oracle_params(cfg), solver_bodies=cfg.solver_bodies)
        self.simulate_calls = 0

    def simulate(self, apply_wrench=False):
        s, N = self.sim_t, self.N
        f = lambda t: t.detach().numpy().astype(np.float64)
        ds = s["dof_state"].view(N, self.nd, 2)
        root, q, qd = f(s["root_states"]), f(ds[:, :, 0]), f(ds[:, :, 1])
        F = f(s["rb_force"]).reshape(N, self.nb, 3) if apply_wrench else None
        T = f(s["rb_torque"]).reshape(N, self.nb, 3) if apply_wrench else None
        for _ in range(self.cfg.substeps):  # wrenches act over the whole simulate() call
            root, q, qd, cf, _d = self.oracle.substep(root, q, qd, f(s["dof_actuation_force"]).reshape(N, self.nd),
                                                     f(s["dof_damping"]), f(s["dof_armature"]), f(s["body_mass_scale"]),
                                                     rb_force=F, rb_torque=T)
        s["root_states"].copy_(torch.tensor(root, dtype=torch.float32))
        ds[:, :, 0] = torch.tensor(q, dtype=torch.float32)
        ds[:, :, 1] = torch.tensor(qd, dtype=torch.float32)
        s["net_contact_force"].copy_(torch.tensor(cf, dtype=torch.float32).reshape(-1, 3))
        self.simulate_calls += 1

    def refresh_rigid_body_state(self):
        pass

    def refresh_dof_force(self, out):
        s, N = self.sim_t, self.N
        ds = s["dof_state"].view(N, self.nd, 2)
        stiff = torch.tensor(self.tables.dof_stiffness, dtype=torch.float32)
        out.copy_((s["dof_actuation_force"].view(N, self.nd) - s["dof_damping"] * ds[:, :, 1] - stiff * ds[:, :, 0]).reshape(-1))

    def refresh_force_sensors(self, sensor_body, sensor_pose, out):
        # (sensor frames: the stand-in keeps no body rotations; world-frame forces are enough for the API surface)
        cf = self.sim_t["net_contact_force"].view(self.N, self.nb, 3)
        out.view(self.N, -1, 6)[:, :, 0:3] = cf[:, sensor_body.long()]
        out.view(self.N, -1, 6)[:, :, 3:6] = 0

    def set_state_indexed(self, ids32, count):
        assert ids32.dtype == torch.int32 and 0 <= count <= ids32.numel()

    def close(self):
        pass
