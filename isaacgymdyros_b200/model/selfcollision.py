"""Self-collision tables of an articulation: what the reference gets from PhysX by creating the actor with
collision filter 0 (`create_actor(env, asset, pose, "humanoid", i, 0, 0)`, tasks/dyros_dynamic_walk.py:354;
docs/_sources/programming/assets.rst.txt:107-109: filter 0 = all shapes of the actor collide with each other), and what
makes `collision_true` (T:590, T:937) fire when an arm hits the torso or a knee hits the other leg.

Model (stated once more, independently, in oracle/selfcollision_oracle.py):
  * every collision primitive of the MJCF (box or cylinder; XML:99-353; capsules and spheres for the stock Humanoid,
    assets/mjcf/nv_humanoid.xml) is kept as an EXACT shape in its link's frame (box: centre, axes, half extents;
    cylinder / capsule: centre, axis, radius, half height; a sphere is a capsule of half height 0) plus a few SAMPLE
    SPHERES (box: its 8 corners, radius 0; cylinder: 3 spheres of radius min(r, h) on its axis; capsule: spheres of its
    own radius on its axis, at most one radius apart, at most 8);
  * two shapes A, B touch when a sample sphere of one penetrates the exact shape of the other: depth = rho - sdf_B(c) > 0;
    the contact pushes the bodies apart along the gradient of sdf_B with the penalty stiffness of the ground contacts, and
    both directions (samples of A in B, samples of B in A) are evaluated;
  * links joined by a joint never collide (as PhysX articulations), and neither do link pairs whose shapes already
    interpenetrate in a rest pose of the robot (geometry that overlaps by design around compound joints: hip yaw / roll
    / pitch links, the wrists) -- decided here, once, with the same narrow phase, for the zero pose and the task's
    initial pose.
  * broad phase per env: bounding spheres of the links (centre and radius over the link's shapes).
The forces only ever have to flag a contact: a non-foot body with |F| > 1 N ends the episode in the same step (T:590).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

KIND_BOX, KIND_CYL, KIND_CAP = 0, 1, 2
MAX_SHAPE_SAMPLES = 8   # csrc/internal.h SC_MAX_SHAPE_SAMPLES


@dataclass
class SelfCollisionTables:
    shape_kind: np.ndarray    # (ns) int32
    shape_link: np.ndarray    # (ns) int32
    shape_body: np.ndarray    # (ns) int32
    shape_sample0: np.ndarray  # (ns+1) int32: samples [sample0[s], sample0[s+1])
    shape_center: np.ndarray  # (ns,3) link frame
    shape_rot: np.ndarray     # (ns,9) row-major, columns = the shape's axes in the link frame (cylinder: column 2 = axis)
    shape_size: np.ndarray    # (ns,3) box: half extents; cylinder: (radius, half height, 0)
    sample: np.ndarray        # (nsamp,4) link frame position, radius
    link_shape0: np.ndarray   # (nl+1) int32: shapes of link l are [link_shape0[l], link_shape0[l+1])
    link_sphere: np.ndarray   # (nl,4) bounding sphere of the link's shapes: centre (link frame), radius (0: no shapes)
    pairs: np.ndarray         # (np,2) int32 candidate link pairs i < j

    @property
    def num_shapes(self) -> int:
        return len(self.shape_kind)


def _boxes_of_body(P: np.ndarray) -> List[tuple]:
    """8 corner points (any order) -> (centre, axes as columns, half extents): the three edges at corner 0 are the nearest
    other corner, then the nearest one perpendicular to it, then the nearest one perpendicular to both."""
    assert P.shape == (8, 3)
    c = P.mean(0)
    D = P - P[0]
    order = np.argsort(np.linalg.norm(D, axis=1))[1:]
    edges = []
    for k in order:
        e = D[k]
        if all(abs(e @ f) < 1e-9 * (np.linalg.norm(e) * np.linalg.norm(f) + 1e-30) + 1e-12 for f in edges):
            edges.append(e)
        if len(edges) == 3:
            break
    if len(edges) != 3:
        raise ValueError("8 points that are not the corners of a box")
    E = np.stack(edges, 1)  # columns: the three edges at corner 0
    half = np.linalg.norm(E, axis=0) / 2
    R = E / (2 * half)
    G = R.T @ R
    if not np.allclose(G, np.eye(3), atol=1e-6):
        raise ValueError("8 points that are not the corners of a box")
    if np.linalg.det(R) < 0:
        R[:, 2] = -R[:, 2]
    return c, R, half


def shapes_from_tables(t) -> List[dict]:
    shapes = []
    rad0 = np.asarray(t.pt_radius) == 0.0
    for b in range(t.num_bodies):
        idx = [i for i in range(len(t.pt_body)) if int(t.pt_body[i]) == b and rad0[i]]
        if len(idx) % 8:
            raise ValueError(f"body {b}: {len(idx)} box-corner points")
        # boxes of one body come in consecutive groups of 8; the solver links' points are sorted by height instead,
        # which is still one box per body for the supported assets (checked by _boxes_of_body)
        for g in range(0, len(idx), 8):
            c, R, half = _boxes_of_body(np.asarray(t.pt_pos)[idx[g:g + 8]])
            samples = np.concatenate([np.asarray(t.pt_pos)[idx[g:g + 8]], np.zeros((8, 1))], 1)
            shapes.append(dict(kind=KIND_BOX, link=int(t.pt_link[idx[g]]), body=b, center=c, rot=R, size=half, samples=samples))
    for k in range(len(t.cyl_link)):
        a = np.asarray(t.cyl_axis[k], float)
        a = a / np.linalg.norm(a)
        u = np.cross(a, [1.0, 0, 0]) if abs(a[0]) < 0.9 else np.cross(a, [0, 1.0, 0])
        u /= np.linalg.norm(u)
        R = np.stack([u, np.cross(a, u), a], 1)
        r, h = float(t.cyl_size[k][0]), float(t.cyl_size[k][1])
        rho = min(r, h)
        c = np.asarray(t.cyl_center[k], float)
        samples = np.array([[*(c + s * (h - rho) * a), rho] for s in (-1.0, 0.0, 1.0)])
        shapes.append(dict(kind=KIND_CYL, link=int(t.cyl_link[k]), body=int(t.cyl_body[k]), center=c, rot=R,
                           size=np.array([r, h, 0.0]), samples=samples))
    for k in range(len(getattr(t, "cap_link", []))):
        a, b = np.asarray(t.cap_ends[k][:3], float), np.asarray(t.cap_ends[k][3:], float)
        r, c, h = float(t.cap_radius[k]), (a + b) / 2, float(np.linalg.norm(b - a)) / 2
        ax = (b - a) / (2 * h) if h > 0 else np.array([0.0, 0, 1.0])
        u = np.cross(ax, [1.0, 0, 0]) if abs(ax[0]) < 0.9 else np.cross(ax, [0, 1.0, 0])
        u /= np.linalg.norm(u)
        R = np.stack([u, np.cross(ax, u), ax], 1)
        n = min(MAX_SHAPE_SAMPLES, 1 + 2 * int(np.ceil(h / r))) if h > 0 else 1
        samples = np.array([[*(c + z * ax), r] for z in (np.linspace(-h, h, n) if n > 1 else [0.0])])
        shapes.append(dict(kind=KIND_CAP, link=int(t.cap_link[k]), body=int(t.cap_body[k]), center=c, rot=R,
                           size=np.array([r, h, 0.0]), samples=samples))
    shapes.sort(key=lambda s: (s["link"], s["kind"], s["body"]))
    return shapes


def sdf(kind: int, size: np.ndarray, x: np.ndarray):
    """Signed distance of points x (..., 3), given in the shape's own frame, and its gradient (unit, shape frame)."""
    if kind == KIND_BOX:
        q = np.abs(x) - size
        outside = np.maximum(q, 0.0)
        dist_out = np.linalg.norm(outside, axis=-1)
        inside = np.minimum(q.max(-1), 0.0)
        d = dist_out + inside
        g_out = outside / np.maximum(dist_out, 1e-30)[..., None]
        ax = np.argmax(q, axis=-1)
        g_in = np.zeros_like(x)
        np.put_along_axis(g_in, ax[..., None], 1.0, -1)
        g = np.where((dist_out > 0)[..., None], g_out, g_in) * np.where(x < 0, -1.0, 1.0)
        return d, g
    r, h = size[0], size[1]
    if kind == KIND_CAP:  # distance to the segment [-h, h] on z, minus the radius
        q = x - np.concatenate([np.zeros_like(x[..., :2]), np.clip(x[..., 2:], -h, h)], -1)
        n = np.linalg.norm(q, axis=-1)
        g = np.where((n > 0)[..., None], q / np.maximum(n, 1e-30)[..., None], np.array([1.0, 0, 0]))
        return n - r, g
    rho = np.linalg.norm(x[..., :2], axis=-1)
    qr, qz = rho - r, np.abs(x[..., 2]) - h
    outside = np.stack([np.maximum(qr, 0.0), np.maximum(qz, 0.0)], -1)
    dist_out = np.linalg.norm(outside, axis=-1)
    d = dist_out + np.minimum(np.maximum(qr, qz), 0.0)
    er = np.concatenate([x[..., :2] / np.maximum(rho, 1e-30)[..., None], np.zeros_like(x[..., :1])], -1)
    ez = np.concatenate([np.zeros_like(x[..., :2]), np.where(x[..., 2:] < 0, -1.0, 1.0)], -1)
    g_out = (outside[..., 0:1] * er + outside[..., 1:2] * ez) / np.maximum(dist_out, 1e-30)[..., None]
    g_in = np.where((qr > qz)[..., None], er, ez)
    return d, np.where((dist_out > 0)[..., None], g_out, g_in)


def link_fk(t, q: np.ndarray):
    """World rotation and origin of every link for joint angles q (nd,), base at the origin with identity rotation."""
    nl = t.num_links
    Rw, pw = [np.eye(3)], [np.zeros(3)]
    for i in range(1, nl):
        p = int(t.link_parent[i])
        a = np.asarray(t.link_axis[i], float)
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        ang = float(q[int(t.link_dof[i])])
        Rj = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
        E = Rj.T @ np.asarray(t.link_E[i], float).reshape(3, 3)
        Rw.append(Rw[p] @ E.T)
        pw.append(pw[p] + Rw[p] @ np.asarray(t.link_r[i], float))
    return Rw, pw


def max_penetration(A: dict, B: dict, RA, pA, RB, pB) -> float:
    """Deepest penetration of A's sample spheres into B (world poses of the two links given)."""
    c = (RA @ A["samples"][:, :3].T).T + pA                  # world
    x = ((c - pB) @ RB - B["center"]) @ B["rot"]              # B's shape frame
    d, _ = sdf(B["kind"], B["size"], x)
    return float((A["samples"][:, 3] - d).max())


def build(t, rest_poses: Sequence[np.ndarray] = (), margin: float = 0.002) -> SelfCollisionTables:
    """`rest_poses`: joint-angle vectors in which touching link pairs are taken to overlap by design (always includes
    the zero pose). `margin`: pairs closer than this in a rest pose are excluded too."""
    shapes = shapes_from_tables(t)
    nl = t.num_links
    by_link = [[s for s in shapes if s["link"] == l] for l in range(nl)]
    poses = [np.zeros(t.num_dofs)] + [np.asarray(q, float) for q in rest_poses]
    fks = [link_fk(t, q) for q in poses]
    parent = np.asarray(t.link_parent)
    pairs = []
    for i in range(nl):
        for j in range(i + 1, nl):
            if not by_link[i] or not by_link[j]:
                continue
            if parent[j] == i or parent[i] == j:
                continue
            touching = False
            for Rw, pw in fks:
                for A in by_link[i]:
                    for B in by_link[j]:
                        if max(max_penetration(A, B, Rw[i], pw[i], Rw[j], pw[j]),
                               max_penetration(B, A, Rw[j], pw[j], Rw[i], pw[i])) > -margin:
                            touching = True
            if not touching:
                pairs.append((i, j))
    link_shape0 = np.zeros(nl + 1, np.int32)
    for l in range(nl):
        link_shape0[l + 1] = link_shape0[l] + len(by_link[l])
    sample0 = np.zeros(len(shapes) + 1, np.int32)
    for k, s in enumerate(shapes):
        sample0[k + 1] = sample0[k] + len(s["samples"])
    link_sphere = np.zeros((nl, 4))
    for l in range(nl):
        if not by_link[l]:
            continue
        pts, rad = [], []
        for s in by_link[l]:
            if s["kind"] == KIND_BOX:
                pts.append(s["samples"][:, :3]); rad.append(np.zeros(8))
            else:
                a, (r, h) = s["rot"][:, 2], s["size"][:2]
                pts.append(np.stack([s["center"] - h * a, s["center"] + h * a])); rad.append(np.full(2, r))
        pts, rad = np.concatenate(pts), np.concatenate(rad)
        c = (pts.min(0) + pts.max(0)) / 2
        link_sphere[l] = [*c, float((np.linalg.norm(pts - c, axis=1) + rad).max())]
    # Order of the candidate pairs = how often their link spheres overlap in poses around the rest poses (most often
    # first; a fixed-seed sample, ties by index). Any order gives the same contacts; this one lets the kernel skip, warp
    # uniformly, the chunks of its shape-pair sweep that hold only pairs whose link spheres are apart (k_self_collision).
    if pairs:
        P = np.array(pairs, np.int32)
        rng = np.random.default_rng(0)
        lo, up = np.asarray(t.dof_lower, float), np.asarray(t.dof_upper, float)
        freq = np.zeros(len(P))
        for base in poses:
            for sigma in (0.05, 0.2, 0.4):
                for _ in range(30):
                    Rw, pw = link_fk(t, np.clip(base + rng.normal(0, sigma, len(base)), lo, up))
                    C = np.array([Rw[l] @ link_sphere[l, :3] + pw[l] for l in range(nl)])
                    d = np.linalg.norm(C[P[:, 0]] - C[P[:, 1]], axis=1)
                    freq += d < link_sphere[P[:, 0], 3] + link_sphere[P[:, 1], 3]
        pairs = [tuple(P[k]) for k in np.argsort(-freq, kind="stable")]
    return SelfCollisionTables(
        shape_kind=np.array([s["kind"] for s in shapes], np.int32), shape_link=np.array([s["link"] for s in shapes], np.int32),
        shape_body=np.array([s["body"] for s in shapes], np.int32), shape_sample0=sample0,
        shape_center=np.array([s["center"] for s in shapes], np.float64).reshape(-1, 3),
        shape_rot=np.array([s["rot"].reshape(9) for s in shapes], np.float64).reshape(-1, 9),
        shape_size=np.array([s["size"] for s in shapes], np.float64).reshape(-1, 3),
        sample=np.concatenate([s["samples"] for s in shapes]).astype(np.float64) if shapes else np.zeros((0, 4)),
        link_shape0=link_shape0, link_sphere=link_sphere, pairs=np.array(pairs, np.int32).reshape(-1, 2))
