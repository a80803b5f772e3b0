"""RobotModel -> flat model tables shared by the CUDA kernels and the CPU oracle.

Fixed-joint children are merged into their parent "link" for dynamics but keep their own row in the
body-indexed API tensors (rigid_body_state, net_contact_force), as the reference's tensors are
body-indexed (docs/_sources/programming/tensors.rst.txt:193,267).

Layout notes (all float arrays are float64 here; the native library down-casts to fp32):
  link l: moving body. link 0 = floating base. link l>=1 has one revolute DOF `link_dof[l]`.
  link_E[l]  3x3 row-major: rotation taking parent-link coordinates to link coordinates at q=0.
  link_r[l]  origin of link l expressed in the parent link frame.
  body_inertia[b] = (m, m*c[3], Ibar_xx, yy, zz, xy, xz, yz) about the *link* origin, link frame.
  contact candidates: points (box corners, sphere/capsule centres with radius) and cylinders.
  sched[T][G]: link processed by lane g at slot t of the branch-parallel recursions (-1 = idle);
      a link's parent is always in an earlier slot. Used by the G-lanes-per-env kernels.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import List

import numpy as np

from .mjcf import RobotModel

LANES_PER_ENV = 4


@dataclass
class ModelTables:
    body_names: List[str]
    dof_names: List[str]
    link_parent: np.ndarray
    link_dof: np.ndarray
    link_E: np.ndarray
    link_r: np.ndarray
    link_axis: np.ndarray
    body_link: np.ndarray
    body_pos: np.ndarray
    body_rot: np.ndarray
    body_inertia: np.ndarray
    dof_lower: np.ndarray
    dof_upper: np.ndarray
    dof_armature: np.ndarray
    dof_damping: np.ndarray
    dof_vel_limit: np.ndarray
    dof_effort: np.ndarray
    pt_link: np.ndarray
    pt_body: np.ndarray
    pt_pos: np.ndarray
    pt_radius: np.ndarray
    cyl_link: np.ndarray
    cyl_body: np.ndarray
    cyl_center: np.ndarray
    cyl_axis: np.ndarray
    cyl_size: np.ndarray  # (radius, half_height)
    solver_links: np.ndarray  # links whose ground contacts go through the PGS solver (feet)
    sched: np.ndarray  # (T, G) int32
    meta: dict = field(default_factory=dict)
    dof_stiffness: np.ndarray = None  # joint spring about q = 0 (MJCF `stiffness`); zeros for TOCABI
    # sphere / capsule primitives as shapes (their end points are also in pt_*): link, body, the two end centres in the
    # link frame (equal for a sphere), radius. Used by the self-collision tables (model/selfcollision.py).
    cap_link: np.ndarray = None
    cap_body: np.ndarray = None
    cap_ends: np.ndarray = None    # (ncap, 6)
    cap_radius: np.ndarray = None

    def __post_init__(self):
        if self.dof_stiffness is None:
            self.dof_stiffness = np.zeros(len(self.dof_names))
        if self.cap_link is None:
            self.cap_link, self.cap_body = np.zeros(0, np.int32), np.zeros(0, np.int32)
            self.cap_ends, self.cap_radius = np.zeros((0, 6)), np.zeros(0)

    @property
    def num_bodies(self) -> int:
        return len(self.body_names)

    @property
    def num_links(self) -> int:
        return int(self.link_parent.shape[0])

    @property
    def num_dofs(self) -> int:
        return len(self.dof_names)

    def total_mass(self) -> float:
        return float(self.body_inertia[:, 0].sum())

    _ARRAYS = ["link_parent", "link_dof", "link_E", "link_r", "link_axis", "body_link", "body_pos", "body_rot",
               "body_inertia", "dof_lower", "dof_upper", "dof_armature", "dof_damping", "dof_vel_limit",
               "dof_effort", "pt_link", "pt_body", "pt_pos", "pt_radius", "cyl_link", "cyl_body", "cyl_center",
               "cyl_axis", "cyl_size", "solver_links", "sched"]

    _OPTIONAL = ["cap_link", "cap_body", "cap_ends", "cap_radius"]

    def save(self, path: str) -> None:
        d = {k: getattr(self, k) for k in self._ARRAYS}
        d["dof_stiffness"] = self.dof_stiffness
        for k in self._OPTIONAL:
            d[k] = getattr(self, k)
        d["names_json"] = np.frombuffer(json.dumps(
            {"body_names": self.body_names, "dof_names": self.dof_names, "meta": self.meta}).encode(), dtype=np.uint8)
        np.savez_compressed(path, **d)

    @classmethod
    def load(cls, path: str) -> "ModelTables":
        z = np.load(path)
        names = json.loads(bytes(z["names_json"]).decode())
        return cls(body_names=names["body_names"], dof_names=names["dof_names"], meta=names.get("meta", {}),
                   dof_stiffness=z["dof_stiffness"] if "dof_stiffness" in z.files else None,
                   **{k: z[k] for k in cls._OPTIONAL if k in z.files},
                   **{k: z[k] for k in cls._ARRAYS})


def branch_schedule(link_parent: np.ndarray, lanes: int = LANES_PER_ENV) -> np.ndarray:
    """Critical-path list scheduling of links 1..nl-1 onto `lanes` lanes.

    Returns sched[T][lanes] with every link exactly once and parent slot < child slot. A lane keeps
    walking down the chain it processed in the previous slot when it can (register-carry friendly).
    """
    nl = len(link_parent)
    children = [[] for _ in range(nl)]
    for l in range(1, nl):
        children[int(link_parent[l])].append(l)
    height = np.zeros(nl, dtype=np.int64)
    for l in range(nl - 1, 0, -1):
        height[l] = 1 + max([height[c] for c in children[l]], default=0)
    done_slot = {0: -1}
    remaining = set(range(1, nl))
    rows = []
    last = [-1] * lanes
    t = 0
    while remaining:
        ready = sorted([l for l in remaining if int(link_parent[l]) in done_slot and done_slot[int(link_parent[l])] < t],
                       key=lambda l: (-height[l], l))
        picked = ready[:lanes]
        row = [-1] * lanes
        unplaced = []
        for l in picked:  # first keep chains on their lane
            p = int(link_parent[l])
            g = last.index(p) if p in last and p != -1 else -1
            if g >= 0 and row[g] == -1:
                row[g] = l
            else:
                unplaced.append(l)
        for l in unplaced:
            row[row.index(-1)] = l
        for g, l in enumerate(row):
            if l >= 0:
                done_slot[l] = t
                remaining.discard(l)
        last = row
        rows.append(row)
        t += 1
    return np.array(rows, dtype=np.int32)


def role_programs(link_parent: np.ndarray, foot_links, roles: int = LANES_PER_ENV) -> np.ndarray:
    """Static partition of the links 1..nl-1 into `roles` programs for the warp-specialised physics kernel.

    The tree is cut into chains (a link's tallest child continues its chain, the other children start new ones);
    chains are dealt to the roles longest-first onto the least loaded role, with every chain that ends in a foot
    (solver) link pinned to a role of its own so that the contact stage of the two feet runs in two different warps.
    Returns prog[T][roles] (-1 padded): column r lists role r's links in ascending (= topological) order. Unlike
    `branch_schedule` there is no slot synchrony: a role waits on per-link flags for what other roles produce."""
    nl = len(link_parent)
    children = [[] for _ in range(nl)]
    for l in range(1, nl):
        children[int(link_parent[l])].append(l)
    height = np.zeros(nl, dtype=np.int64)
    for l in range(nl - 1, 0, -1):
        height[l] = 1 + max([height[c] for c in children[l]], default=0)
    chains, stack = [], list(children[0])
    while stack:
        head = stack.pop()
        chain, l = [], head
        while True:
            chain.append(l)
            if not children[l]:
                break
            nxt = max(children[l], key=lambda c: (height[c], -c))
            stack.extend(c for c in children[l] if c != nxt)
            l = nxt
        chains.append(chain)
    foot_links = [int(f) for f in foot_links]
    foot_chains = [c for c in chains if any(f in c for f in foot_links)]
    other = sorted([c for c in chains if c not in foot_chains], key=lambda c: -len(c))
    if len(foot_chains) > roles:
        raise ValueError("more foot chains than roles")
    load = [0] * roles
    members = [[] for _ in range(roles)]
    # feet take the last roles; the contact stage is charged as extra load so that long arm chains go elsewhere
    for k, c in enumerate(foot_chains):
        r = roles - 1 - k
        members[r] += c
        load[r] += len(c) + 8
    for c in other:
        r = int(np.argmin(load))
        members[r] += c
        load[r] += len(c)
    T = max(len(mm) for mm in members)
    prog = -np.ones((T, roles), dtype=np.int32)
    for r, mm in enumerate(members):
        for k, l in enumerate(sorted(mm)):
            prog[k, r] = l
    return prog


def build_tables(model: RobotModel, *, solver_bodies: List[str], vel_limit: float = 4.03,
                 default_damping: float = 0.0, lanes: int = LANES_PER_ENV) -> ModelTables:
    """Flatten `model`. `solver_bodies`: names of the bodies whose ground contact is constraint-solved
    (for TOCABI: L_Foot_Link, R_Foot_Link; the whole merged link they belong to is included)."""
    nb = model.num_bodies
    body_link = np.zeros(nb, dtype=np.int32)
    body_pos = np.zeros((nb, 3))
    body_rot = np.zeros((nb, 3, 3))
    link_body = []  # representative body of each link
    link_parent, link_E, link_r, link_axis, link_dof = [], [], [], [], []
    dof_names, lo, up, arm, damp, eff = [], [], [], [], [], []
    stiff = []
    for bi, b in enumerate(model.bodies):
        hinges = [j for j in b.joints if j.type == "hinge"]
        frees = [j for j in b.joints if j.type == "free"]
        if b.parent < 0:
            if not frees:
                raise NotImplementedError("fixed-base articulations are not on this path")
            body_link[bi] = 0
            body_pos[bi] = 0.0
            body_rot[bi] = np.eye(3)
            link_body.append(bi)
            link_parent.append(-1)
            link_E.append(np.eye(3))
            link_r.append(np.zeros(3))
            link_axis.append(np.zeros(3))
            link_dof.append(-1)
            continue
        pl = body_link[b.parent]
        # body frame relative to the parent's link frame
        Rp, pp = body_rot[b.parent], body_pos[b.parent]
        R_rel = Rp @ b.rot
        p_rel = pp + Rp @ b.pos
        if hinges:
            # One link per hinge. A body with k hinges (MuJoCo applies them in order, each about its anchor) becomes a
            # chain of k links: the first k-1 are massless, the body itself rides on the last one. Every link frame has
            # the body's orientation at q = 0 and its origin at the hinge's anchor.
            prev_anchor = None
            for t, j in enumerate(hinges):
                l = len(link_parent)
                if t == 0:
                    link_parent.append(int(pl))
                    link_E.append(R_rel.T)
                    link_r.append(p_rel + R_rel @ j.pos)
                else:
                    link_parent.append(l - 1)
                    link_E.append(np.eye(3))
                    link_r.append(j.pos - prev_anchor)
                prev_anchor = j.pos
                link_axis.append(j.axis)
                link_dof.append(len(dof_names))
                link_body.append(bi)
                dof_names.append(j.name)
                lo.append(min(j.range))
                up.append(max(j.range))
                arm.append(j.armature)
                damp.append(j.damping if j.damping else default_damping)
                stiff.append(j.stiffness)
                lim = model.actuators.get(j.name, (-np.inf, np.inf, 1.0))
                eff.append(max(abs(lim[0]), abs(lim[1])) * abs(lim[2]))
            body_link[bi] = len(link_parent) - 1
            body_pos[bi] = -prev_anchor
            body_rot[bi] = np.eye(3)
        else:  # fixed: merge
            body_link[bi] = pl
            body_pos[bi] = p_rel
            body_rot[bi] = R_rel
    nl = len(link_parent)
    body_inertia = np.zeros((nb, 10))
    for bi, b in enumerate(model.bodies):
        R, p = body_rot[bi], body_pos[bi]
        c = R @ b.com + p
        Ic = R @ b.inertia @ R.T
        Io = Ic + b.mass * (c @ c * np.eye(3) - np.outer(c, c))
        body_inertia[bi] = [b.mass, *(b.mass * c), Io[0, 0], Io[1, 1], Io[2, 2], Io[0, 1], Io[0, 2], Io[1, 2]]
    pt_link, pt_body, pt_pos, pt_rad = [], [], [], []
    cyl_link, cyl_body, cyl_c, cyl_a, cyl_s = [], [], [], [], []
    cap_link, cap_body, cap_ends, cap_rad = [], [], [], []
    for bi, b in enumerate(model.bodies):
        R, p = body_rot[bi], body_pos[bi]
        for g in b.geoms:
            Rg = R @ g.rot
            cg = R @ g.pos + p
            if g.type == "box":
                for sx in (-1, 1):
                    for sy in (-1, 1):
                        for sz in (-1, 1):
                            pt_link.append(body_link[bi]); pt_body.append(bi)
                            pt_pos.append(cg + Rg @ (np.array([sx, sy, sz]) * g.size[:3])); pt_rad.append(0.0)
            elif g.type == "sphere":
                pt_link.append(body_link[bi]); pt_body.append(bi); pt_pos.append(cg); pt_rad.append(g.size[0])
                cap_link.append(body_link[bi]); cap_body.append(bi); cap_ends.append([*cg, *cg]); cap_rad.append(g.size[0])
            elif g.type == "capsule":
                ends = [cg + Rg @ np.array([0, 0, s * g.size[1]]) for s in (-1, 1)]
                cap_link.append(body_link[bi]); cap_body.append(bi); cap_ends.append([*ends[0], *ends[1]]); cap_rad.append(g.size[0])
                for s in (-1, 1):
                    pt_link.append(body_link[bi]); pt_body.append(bi)
                    pt_pos.append(cg + Rg @ np.array([0, 0, s * g.size[1]])); pt_rad.append(g.size[0])
            elif g.type == "cylinder":
                cyl_link.append(body_link[bi]); cyl_body.append(bi); cyl_c.append(cg)
                cyl_a.append(Rg[:, 2]); cyl_s.append([g.size[0], g.size[1]])
    # candidates of the solver links first, sole (lowest local z) first: a cap on active points keeps the sole
    sl = sorted({int(body_link[model.body_index(n)]) for n in solver_bodies})
    order = sorted(range(len(pt_link)), key=lambda i: (0 if pt_link[i] in sl else 1, pt_link[i],
                                                       pt_pos[i][2] if pt_link[i] in sl else 0.0, i))
    pt_link = [pt_link[i] for i in order]; pt_body = [pt_body[i] for i in order]
    pt_pos = [pt_pos[i] for i in order]; pt_rad = [pt_rad[i] for i in order]
    link_parent = np.array(link_parent, dtype=np.int32)
    f64 = lambda a, shape: np.array(a, dtype=np.float64).reshape(shape)
    return ModelTables(
        body_names=[b.name for b in model.bodies], dof_names=dof_names,
        link_parent=link_parent, link_dof=np.array(link_dof, dtype=np.int32),
        link_E=f64(link_E, (nl, 9)), link_r=f64(link_r, (nl, 3)), link_axis=f64(link_axis, (nl, 3)),
        body_link=body_link, body_pos=body_pos, body_rot=body_rot.reshape(nb, 9), body_inertia=body_inertia,
        dof_lower=f64(lo, -1), dof_upper=f64(up, -1), dof_armature=f64(arm, -1), dof_damping=f64(damp, -1),
        dof_vel_limit=np.full(len(dof_names), float(vel_limit)), dof_effort=f64(eff, -1), dof_stiffness=f64(stiff, -1),
        pt_link=np.array(pt_link, dtype=np.int32), pt_body=np.array(pt_body, dtype=np.int32),
        pt_pos=f64(pt_pos, (-1, 3)), pt_radius=f64(pt_rad, -1),
        cyl_link=np.array(cyl_link, dtype=np.int32), cyl_body=np.array(cyl_body, dtype=np.int32),
        cyl_center=f64(cyl_c, (-1, 3)), cyl_axis=f64(cyl_a, (-1, 3)), cyl_size=f64(cyl_s, (-1, 2)),
        cap_link=np.array(cap_link, dtype=np.int32), cap_body=np.array(cap_body, dtype=np.int32),
        cap_ends=f64(cap_ends, (-1, 6)), cap_radius=f64(cap_rad, -1),
        solver_links=np.array(sl, dtype=np.int32), sched=branch_schedule(link_parent, lanes),
        meta={"model": model.name, "lanes": lanes})
