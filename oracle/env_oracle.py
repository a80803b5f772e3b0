"""TEST INFRASTRUCTURE ONLY -- the whole env step on the CPU: oracle/task_oracle.py (restatement of the reference's
DyrosDynamicWalk methods, pinned to the reference by golden vectors) around oracle/physics_oracle.py (fp64 statement
of our `simulate` model; PhysX parity unpinned, see its header) in the place of the reference's gym.simulate call
(tasks/dyros_dynamic_walk.py:525). Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg only."""
from __future__ import annotations

import numpy as np

from . import task_oracle as O
from .physics_oracle import PhysicsOracle, PhysParams

f32 = np.float32


class EnvOracle:
    def __init__(self, N, tables, mocap, obs_norm, phys_params: PhysParams = None, task_params: O.Params = None,
                 damping=None, armature=None, mass_scale=None, rng=None, self_collision=True):
        from isaacgymdyros_b200.core import ARMATURE  # constants only
        self.N = N
        self.tables = tables
        self.phys = PhysicsOracle(tables, phys_params or PhysParams())
        self.sc = None
        if self_collision:  # the actor's shapes collide with each other (create_actor filter 0, T:354)
            from isaacgymdyros_b200.core import self_collision_tables  # tables only
            self.sc = self_collision_tables(tables)
        self.damping = np.full((N, 33), 0.1) if damping is None else np.asarray(damping, float)
        self.armature = np.tile(np.array(ARMATURE), (N, 1)) if armature is None else np.asarray(armature, float)
        self.mass_scale = np.ones((N, 38)) if mass_scale is None else np.asarray(mass_scale, float)
        total_mass = (tables.body_inertia[:, 0][None, :] * self.mass_scale).sum(1).astype(f32)
        self.s, self.c = O.new_state(N, mocap, obs_norm, total_mass, tables.dof_lower, tables.dof_upper,
                                     task_params or O.Params(), rng=rng)

    def simulate(self, s, tau, ext):
        """gym.set_dof_actuation_force_tensor + gym.simulate + refresh (T:520-526) on the oracle state dict."""
        if self.sc is not None:  # from the poses the sub-step starts from, as its ground contact forces are
            from .selfcollision_oracle import self_contact_forces_batch
            _, Rw, pw = self.phys.kinematics(s["root_states"].astype(float), s["dof_pos"].astype(float))
            sc_f = self_contact_forces_batch(self.sc, Rw, pw, self.phys.p.pen_k, self.phys.p.pen_fmax)
        root, q, qd, cf, _ = self.phys.substep(s["root_states"].astype(float), s["dof_pos"].astype(float),
                                               s["dof_vel"].astype(float), tau.astype(float), self.damping,
                                               self.armature, self.mass_scale,
                                               push=None if ext is None else ext.astype(float))
        s["root_states"], s["dof_pos"], s["dof_vel"] = root.astype(f32), q.astype(f32), qd.astype(f32)
        if self.sc is not None:
            cf = cf + sc_f[:, :cf.shape[1]]
        s["contact_forces"] = cf.astype(f32)

    def step(self, actions, noise):
        return O.step(self.s, self.c, actions, noise, self.simulate)
