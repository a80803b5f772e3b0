"""TEST INFRASTRUCTURE ONLY -- CPU (numpy, float64, dense) statement of the `gym.simulate` model that the
CUDA physics kernel implements (reference call site: tasks/dyros_dynamic_walk.py:525; the reference's
own implementation is the closed-source PhysX inside Isaac Gym Preview 4, python/setup.py:32, whose
binaries are absent from /root/reference: .MISSING_LARGE_BLOBS:12-30).

PARITY UNPINNED against PhysX: no golden trajectory, fixture or test of the reference pins results at the
`simulate` boundary, and PhysX cannot run here. This oracle therefore pins OUR model:
it follows the reference's model *parameters* (MJCF tree/inertias dyros_tocabi.xml:95-367, dt/solver
parameters DyrosDynamicWalk.yaml:37-56, dof properties dyros_dynamic_walk.py:363-373) and is validated by
physical invariants in tests/test_physics_oracle.py (free fall, momentum, energy, weight on the ground).

It is deliberately a *different algorithm* from the kernel: joint-space (link Jacobians, dense mass
matrix, dense solves) instead of the kernel's O(n) articulated-body recursions, so agreement between the
two checks the kernel's recursions rather than restating them. Batched over envs (leading axis N).

Conventions: spatial motion [w; v], force [n; f], link coordinates; X_i maps parent motion to link i;
quaternions xyzw (Isaac Gym, docs/_sources/programming/tensors.rst.txt:50-62).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

MAX_ACTIVE_PTS = 4


@dataclass
class PhysParams:
    dt: float = 0.002
    gravity: tuple = (0.0, 0.0, -9.81)
    contact_offset: float = 0.002
    max_depen_vel: float = 10.0
    erp: float = 0.2
    mu: float = 1.0
    pen_k: float = 2.0e5
    pen_c: float = 2.0e3
    pen_fmax: float = 2.0e4
    max_ang_vel: float = 100.0
    sweeps: int = 5
    clamp_effort: bool = False
    vel_limit: float = 4.03


def skew(v):
    z = np.zeros(v.shape[:-1])
    return np.stack([np.stack([z, -v[..., 2], v[..., 1]], -1), np.stack([v[..., 2], z, -v[..., 0]], -1),
                     np.stack([-v[..., 1], v[..., 0], z], -1)], -2)


def quat_to_mat(q):
    """xyzw -> rotation (body to world)."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
                     np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
                     np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], -2)


def axis_rot(axis, q):
    """Rodrigues: rotation by angle q (N,) about unit `axis` (3,)."""
    K = skew(np.asarray(axis, float))
    c, s = np.cos(q)[:, None, None], np.sin(q)[:, None, None]
    return np.eye(3) + s * K + (1 - c) * (K @ K)


def spatial_inertia(par):
    """(..., 10) [m, h3, Ixx, Iyy, Izz, Ixy, Ixz, Iyz] about the link origin -> (..., 6, 6)."""
    m, h = par[..., 0], par[..., 1:4]
    I = np.zeros(par.shape[:-1] + (6, 6))
    I[..., 0, 0], I[..., 1, 1], I[..., 2, 2] = par[..., 4], par[..., 5], par[..., 6]
    I[..., 0, 1] = I[..., 1, 0] = par[..., 7]
    I[..., 0, 2] = I[..., 2, 0] = par[..., 8]
    I[..., 1, 2] = I[..., 2, 1] = par[..., 9]
    H = skew(h)
    I[..., 0:3, 3:6] = H
    I[..., 3:6, 0:3] = np.swapaxes(H, -1, -2)
    for k in range(3):
        I[..., 3 + k, 3 + k] = m
    return I


def crf_apply(v, f):
    """v x* f."""
    w, vl, n, fl = v[..., :3], v[..., 3:], f[..., :3], f[..., 3:]
    return np.concatenate([np.cross(w, n) + np.cross(vl, fl), np.cross(w, fl)], -1)


def crm_apply(v, m):
    """v x m."""
    w, vl = v[..., :3], v[..., 3:]
    return np.concatenate([np.cross(w, m[..., :3]), np.cross(w, m[..., 3:]) + np.cross(vl, m[..., :3])], -1)


class PhysicsOracle:
    def __init__(self, tables, params: PhysParams = None, solver_bodies=("L_Foot_Link", "R_Foot_Link")):
        self.t = tables
        self.p = params or PhysParams()
        t = tables
        self.nl, self.nb, self.nd = t.num_links, t.num_bodies, t.num_dofs
        solver_ids = [t.body_names.index(n) for n in solver_bodies]
        self.pt_solver = np.isin(t.pt_body, solver_ids)
        self.feet = sorted(set(int(l) for l in t.pt_link[self.pt_solver]))
        self.foot_pts = {f: [i for i in range(len(t.pt_link)) if self.pt_solver[i] and t.pt_link[i] == f] for f in self.feet}
        self.link_bodies = [[b for b in range(self.nb) if t.body_link[b] == l] for l in range(self.nl)]
        self.link_pts = [[i for i in range(len(t.pt_link)) if t.pt_link[i] == l and not self.pt_solver[i]] for l in range(self.nl)]
        self.link_cyls = [[i for i in range(len(t.cyl_link)) if t.cyl_link[i] == l] for l in range(self.nl)]

    # ------------------------------------------------------------------ kinematics
    def kinematics(self, root, q):
        t, nl = self.t, self.nl
        N = root.shape[0]
        Rw = [quat_to_mat(root[:, 3:7])]
        pw = [root[:, 0:3].copy()]
        X = [None]
        for i in range(1, nl):
            p = int(t.link_parent[i])
            E = np.swapaxes(axis_rot(t.link_axis[i], q[:, t.link_dof[i]]), -1, -2) @ t.link_E[i].reshape(3, 3)
            r = t.link_r[i]
            Xi = np.zeros((N, 6, 6))
            Xi[:, :3, :3] = E
            Xi[:, 3:, 3:] = E
            Xi[:, 3:, :3] = -E @ skew(r)
            X.append(Xi)
            Rw.append(Rw[p] @ np.swapaxes(E, -1, -2))
            pw.append(pw[p] + Rw[p] @ r)
        return X, Rw, pw

    def jacobians(self, X):
        t, nl, nd = self.t, self.nl, self.nd
        N = X[1].shape[0]
        J0 = np.zeros((N, 6, 6 + nd))
        J0[:, :, :6] = np.eye(6)
        J = [J0]
        for i in range(1, nl):
            Ji = X[i] @ J[int(t.link_parent[i])]
            Ji[:, :3, 6 + t.link_dof[i]] += t.link_axis[i]
            J.append(Ji)
        return J

    def link_inertias(self, mass_scale):
        """(N, nb) scale -> list over links of (N,6,6)."""
        par = self.t.body_inertia[None, :, :] * mass_scale[:, :, None]
        return [spatial_inertia(par[:, bs, :].sum(1)) for bs in self.link_bodies]

    # ------------------------------------------------------------------ one sub-step
    def substep(self, root, q, qd, tau, damping, armature, mass_scale, push=None, rb_force=None, rb_torque=None):
        """Advance one sub-step. Arrays float64: root (N,13), q/qd/tau/damping/armature (N,nd), mass_scale (N,nb);
        push (N,3) world force at the base body's COM; rb_force/rb_torque (N,nb,3) world wrench at body COMs.
        Returns root', q', qd', net_contact_force (N,nb,3) and a dict of diagnostics."""
        t, p, nl, nd, nb = self.t, self.p, self.nl, self.nd, self.nb
        N = root.shape[0]
        dt = p.dt
        g = np.asarray(p.gravity, float)
        X, Rw, pw = self.kinematics(root, q)
        J = self.jacobians(X)
        I = self.link_inertias(mass_scale)
        R0 = Rw[0]
        v0 = np.concatenate([np.einsum("nji,nj->ni", R0, root[:, 10:13]), np.einsum("nji,nj->ni", R0, root[:, 7:10])], -1)
        vgen = np.concatenate([v0, qd], -1)
        v = [np.einsum("nij,nj->ni", Ji, vgen) for Ji in J]
        # velocity-product accelerations
        avp = [np.zeros((N, 6))]
        for i in range(1, nl):
            S = np.zeros(6)
            S[:3] = t.link_axis[i]
            c = crm_apply(v[i], S[None, :] * qd[:, t.link_dof[i], None])
            avp.append(np.einsum("nij,nj->ni", X[i], avp[int(t.link_parent[i])]) + c)
        contact = np.zeros((N, nb, 3))
        M = np.zeros((N, 6 + nd, 6 + nd))
        h = np.zeros((N, 6 + nd))
        par = t.body_inertia[None, :, :] * mass_scale[:, :, None]
        for i in range(nl):
            Ii = I[i]
            M += np.swapaxes(J[i], -1, -2) @ Ii @ J[i]
            ag = np.concatenate([np.zeros((N, 3)), np.einsum("nji,j->ni", Rw[i], g)], -1)
            fi = np.einsum("nij,nj->ni", Ii, avp[i] - ag) + crf_apply(v[i], np.einsum("nij,nj->ni", Ii, v[i]))
            fi = fi - self._external_wrench(i, Rw[i], pw[i], v[i], contact, par, push, rb_force, rb_torque)
            h += np.einsum("nji,nj->ni", J[i], fi)
        idx = np.arange(nd)
        stiff = np.asarray(getattr(t, "dof_stiffness", np.zeros(nd)), float)  # joint spring about q = 0, implicit
        M[:, 6 + idx, 6 + idx] += armature + dt * damping + dt * dt * stiff
        tq = tau.copy()
        if p.clamp_effort:
            tq = np.clip(tq, -t.dof_effort, t.dof_effort)
        rhs = -h
        rhs[:, 6:] += tq - damping * qd - stiff * (q + dt * qd)
        Minv = np.linalg.inv(M)
        acc = np.einsum("nij,nj->ni", Minv, rhs)
        vstar = vgen + dt * acc
        # ---- constraint-solved ground contact of the solver points (block-Jacobi between feet, Gauss-Seidel inside)
        feet = self.feet
        nF = len(feet)
        Jf = [J[f] for f in feet]
        Om = [[Jf[a] @ Minv @ np.swapaxes(Jf[b], -1, -2) for b in range(nF)] for a in range(nF)]
        V = [np.einsum("nij,nj->ni", Jf[a], vstar) for a in range(nF)]
        P = [np.zeros((N, 6)) for _ in range(nF)]
        rows = []  # per foot: list of (Jrow (N,3,6), bias (N,), active (N,), body)
        for a, f in enumerate(feet):
            cnt = np.zeros(N, int)
            pts = []
            for i in self.foot_pts[f]:
                x, rad = t.pt_pos[i], t.pt_radius[i]
                z = pw[f][:, 2] + np.einsum("nj,j->n", Rw[f][:, 2, :], x)
                phi = z - rad
                act = (phi < p.contact_offset) & (cnt < MAX_ACTIVE_PTS)
                cnt += act
                xs = x[None, :] - rad * Rw[f][:, 2, :]  # surface point, link coords (normal_l = row 2 of Rw)
                dirs = np.stack([Rw[f][:, 2, :], Rw[f][:, 0, :], Rw[f][:, 1, :]], 1)  # n, t1 (world x), t2 (world y)
                Jr = np.concatenate([np.cross(xs[:, None, :], dirs), dirs], -1)  # (N,3,6)
                bias = np.where(phi >= 0, -phi / dt, np.minimum(-p.erp * phi / dt, p.max_depen_vel))
                pts.append((Jr, bias, act, int(t.pt_body[i])))
            rows.append(pts)
        lam = [[np.zeros((N, 3)) for _ in rows[a]] for a in range(nF)]
        for _ in range(p.sweeps):
            dP = [np.zeros((N, 6)) for _ in range(nF)]
            for a in range(nF):
                Oaa = Om[a][a]
                for k, (Jr, bias, act, _b) in enumerate(rows[a]):
                    for d in range(3):
                        Jd = Jr[:, d, :]
                        cvec = np.einsum("nij,nj->ni", Oaa, Jd)
                        w = np.einsum("ni,ni->n", Jd, cvec)
                        vrel = np.einsum("ni,ni->n", Jd, V[a])
                        if d == 0:
                            new = np.maximum(lam[a][k][:, 0] + (bias - vrel) / w, 0.0)
                        else:
                            lim = p.mu * lam[a][k][:, 0]
                            new = np.clip(lam[a][k][:, d] - vrel / w, -lim, lim)
                        delta = np.where(act, new - lam[a][k][:, d], 0.0)
                        lam[a][k][:, d] += delta
                        V[a] = V[a] + cvec * delta[:, None]
                        dP[a] = dP[a] + Jd * delta[:, None]
            for a in range(nF):
                for b in range(nF):
                    if b != a:
                        V[a] = V[a] + np.einsum("nij,nj->ni", Om[a][b], dP[b])
                P[a] = P[a] + dP[a]
        imp = np.zeros((N, 6 + nd))
        for a in range(nF):
            imp += np.einsum("nji,nj->ni", Jf[a], P[a])
            for k, (_Jr, _bias, _act, b) in enumerate(rows[a]):
                contact[:, b, 0] += lam[a][k][:, 1] / dt
                contact[:, b, 1] += lam[a][k][:, 2] / dt
                contact[:, b, 2] += lam[a][k][:, 0] / dt
        vnew = vstar + np.einsum("nij,nj->ni", Minv, imp)
        # ---- joints: velocity cap, integrate, limit projection
        qdn = np.clip(vnew[:, 6:], -p.vel_limit, p.vel_limit)
        qn = q + dt * qdn
        over, under = qn > t.dof_upper, qn < t.dof_lower
        qdn = np.where(over, np.minimum(qdn, 0.0), np.where(under, np.maximum(qdn, 0.0), qdn))
        qn = np.clip(qn, t.dof_lower, t.dof_upper)
        # ---- base
        wb, vb = vnew[:, :3], vnew[:, 3:6] + dt * np.cross(v0[:, :3], v0[:, 3:])
        ww = np.einsum("nij,nj->ni", R0, wb)
        vw = np.einsum("nij,nj->ni", R0, vb)
        wn = np.linalg.norm(ww, axis=-1, keepdims=True)
        ww = np.where(wn > p.max_ang_vel, ww * (p.max_ang_vel / np.maximum(wn, 1e-30)), ww)
        rootn = root.copy()
        rootn[:, 0:3] = root[:, 0:3] + dt * vw
        qx = root[:, 3:7]
        dq = 0.5 * dt * np.stack([ww[:, 0] * qx[:, 3] + ww[:, 1] * qx[:, 2] - ww[:, 2] * qx[:, 1],
                                  -ww[:, 0] * qx[:, 2] + ww[:, 1] * qx[:, 3] + ww[:, 2] * qx[:, 0],
                                  ww[:, 0] * qx[:, 1] - ww[:, 1] * qx[:, 0] + ww[:, 2] * qx[:, 3],
                                  -ww[:, 0] * qx[:, 0] - ww[:, 1] * qx[:, 1] - ww[:, 2] * qx[:, 2]], -1)
        qq = qx + dq
        rootn[:, 3:7] = qq / np.linalg.norm(qq, axis=-1, keepdims=True)
        rootn[:, 7:10] = vw
        rootn[:, 10:13] = ww
        diag = {"M": M, "acc": acc, "vstar": vstar, "lam": lam, "J": J, "I": I, "v": v, "Rw": Rw, "pw": pw}
        return rootn, qn, qdn, contact, diag

    def _external_wrench(self, i, Rw, pw, v, contact, par, push, rb_force, rb_torque):
        """Penalty ground contact of link i's non-solved candidates + applied body wrenches, as a spatial
        force in link coordinates (N,6). Adds world contact forces into `contact` per body."""
        t, p = self.t, self.p
        N = Rw.shape[0]
        out = np.zeros((N, 6))

        def add_point(xs_l, depth, body):
            """xs_l (N,3) contact location in link coords, depth (N,) > 0 where penetrating."""
            vel_l = v[:, 3:] + np.cross(v[:, :3], xs_l)
            vel_w = np.einsum("nij,nj->ni", Rw, vel_l)
            on = depth > 0
            fn = np.clip(p.pen_k * depth - p.pen_c * vel_w[:, 2], 0.0, p.pen_fmax)
            speed = np.sqrt(vel_w[:, 0] ** 2 + vel_w[:, 1] ** 2)
            coef = np.minimum(p.pen_c, p.mu * fn / np.maximum(speed, 1e-6))
            Fw = np.stack([-coef * vel_w[:, 0], -coef * vel_w[:, 1], fn], -1) * on[:, None]
            contact[:, body, :] += Fw
            fl = np.einsum("nji,nj->ni", Rw, Fw)
            out[:, :3] += np.cross(xs_l, fl)
            out[:, 3:] += fl

        nrm_l = Rw[:, 2, :]  # world z in link coords
        for k in self.link_pts[i]:
            x, rad = t.pt_pos[k], t.pt_radius[k]
            z = pw[:, 2] + np.einsum("nj,j->n", nrm_l, x)
            add_point(x[None, :] - rad * nrm_l, rad - z, int(t.pt_body[k]))
        for k in self.link_cyls[i]:
            c, a, (rad, hh) = t.cyl_center[k], t.cyl_axis[k], t.cyl_size[k]
            az = np.einsum("nj,j->n", nrm_l, a)  # world z component of the axis
            s = np.where(az >= 0, -1.0, 1.0)
            d_l = -(nrm_l - az[:, None] * a[None, :])  # downward radial direction, link coords
            dn = np.linalg.norm(d_l, axis=-1)
            rim = c[None, :] + (s * hh)[:, None] * a[None, :] + np.where(dn[:, None] > 1e-6, rad * d_l / np.maximum(dn, 1e-30)[:, None], 0.0)
            z = pw[:, 2] + np.einsum("nj,nj->n", nrm_l, rim)
            add_point(rim, -z, int(t.cyl_body[k]))
        for b in self.link_bodies[i]:
            F = np.zeros((N, 3))
            Tq = np.zeros((N, 3))
            if push is not None and b == 0:
                F = F + push
            if rb_force is not None:
                F = F + rb_force[:, b, :]
            if rb_torque is not None:
                Tq = Tq + rb_torque[:, b, :]
            if push is None and rb_force is None and rb_torque is None:
                continue
            com = par[:, b, 1:4] / np.maximum(par[:, b, 0:1], 1e-30)
            fl = np.einsum("nji,nj->ni", Rw, F)
            out[:, :3] += np.cross(com, fl) + np.einsum("nji,nj->ni", Rw, Tq)
            out[:, 3:] += fl
        return out

    # ------------------------------------------------------------------ rigid body state (docs tensors.rst.txt:193-207)
    def rigid_body_state(self, root, q, qd):
        t = self.t
        N = root.shape[0]
        X, Rw, pw = self.kinematics(root, q)
        J = self.jacobians(X)
        R0 = Rw[0]
        v0 = np.concatenate([np.einsum("nji,nj->ni", R0, root[:, 10:13]), np.einsum("nji,nj->ni", R0, root[:, 7:10])], -1)
        vgen = np.concatenate([v0, qd], -1)
        out = np.zeros((N, self.nb, 13))
        for b in range(self.nb):
            l = int(t.body_link[b])
            v = np.einsum("nij,nj->ni", J[l], vgen)
            Rb = Rw[l] @ t.body_rot[b].reshape(3, 3)
            out[:, b, 0:3] = pw[l] + Rw[l] @ t.body_pos[b]
            out[:, b, 3:7] = mat_to_quat(Rb)
            out[:, b, 7:10] = np.einsum("nij,nj->ni", Rw[l], v[:, 3:] + np.cross(v[:, :3], t.body_pos[b][None, :]))
            out[:, b, 10:13] = np.einsum("nij,nj->ni", Rw[l], v[:, :3])
        return out

    # ------------------------------------------------------------------ diagnostics used by the invariant tests
    def momentum_energy(self, root, q, qd, mass_scale):
        """World-frame linear momentum, angular momentum about the world origin, kinetic and potential energy."""
        X, Rw, pw = self.kinematics(root, q)
        J = self.jacobians(X)
        I = self.link_inertias(mass_scale)
        N = root.shape[0]
        R0 = Rw[0]
        v0 = np.concatenate([np.einsum("nji,nj->ni", R0, root[:, 10:13]), np.einsum("nji,nj->ni", R0, root[:, 7:10])], -1)
        vgen = np.concatenate([v0, qd], -1)
        par = self.t.body_inertia[None, :, :] * mass_scale[:, :, None]
        lin, ang, ke, pe = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros(N), np.zeros(N)
        g = np.asarray(self.p.gravity, float)
        for i in range(self.nl):
            v = np.einsum("nij,nj->ni", J[i], vgen)
            hmom = np.einsum("nij,nj->ni", I[i], v)
            ke += 0.5 * np.einsum("ni,ni->n", v, hmom)
            f_w = np.einsum("nij,nj->ni", Rw[i], hmom[:, 3:])
            n_w = np.einsum("nij,nj->ni", Rw[i], hmom[:, :3]) + np.cross(pw[i], f_w)
            lin += f_w
            ang += n_w
            pl = par[:, self.link_bodies[i], :].sum(1)
            com_w = pw[i] * pl[:, 0:1] + np.einsum("nij,nj->ni", Rw[i], pl[:, 1:4])
            pe += -np.einsum("ni,i->n", com_w, g)
        return lin, ang, ke, pe


def mat_to_quat(R):
    """(N,3,3) -> xyzw, w >= 0 branch-free enough for tests."""
    N = R.shape[0]
    q = np.zeros((N, 4))
    for n in range(N):
        m = R[n]
        tr = m[0, 0] + m[1, 1] + m[2, 2]
        if tr > 0:
            s = np.sqrt(tr + 1.0) * 2
            q[n] = [(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s]
        elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
            s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
            q[n] = [0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s, (m[2, 1] - m[1, 2]) / s]
        elif m[1, 1] > m[2, 2]:
            s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
            q[n] = [(m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s, (m[0, 2] - m[2, 0]) / s]
        else:
            s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
            q[n] = [(m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s, (m[1, 0] - m[0, 1]) / s]
    return q
