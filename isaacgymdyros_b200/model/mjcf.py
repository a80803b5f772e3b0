"""MJCF -> RobotModel loader (host side, one-off at `gym.load_asset`).

Replaces the closed-source MJCF importer inside `libcarb.gym.plugin.so` for the asset
classes this path needs (reference call site: tasks/dyros_dynamic_walk.py:293
`gym.load_asset(sim, asset_root, asset_file, asset_options)`; model source of truth:
assets/mjcf/dyros_tocabi/xml/dyros_tocabi.xml:95-410).

Pure Python (`xml.etree`), no MuJoCo. Supported subset: nested <body>, <joint type=free|hinge>,
<inertial fullinertia|diaginertia [quat]>, <geom type=box|cylinder|capsule|sphere> with
pos/quat/size/fromto, <default class=...> inheritance, <compiler angle=...>, <actuator><motor>.
Mesh geoms are visual only in the reference model (contype=conaffinity=0) and are skipped.

Conventions: quaternions here are MJCF order (w, x, y, z); rotation matrices map child-frame
coordinates to parent-frame coordinates (p_parent = R @ p_child + pos).
"""
from __future__ import annotations

import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np


def quat_wxyz_to_mat(q) -> np.ndarray:
    w, x, y, z = [float(v) for v in q]
    n = math.sqrt(w * w + x * x + y * y + z * z)
    if n == 0.0:
        return np.eye(3)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ], dtype=np.float64)


def euler_xyz_to_mat(e) -> np.ndarray:
    """MuJoCo default eulerseq 'xyz' (intrinsic rotations about x, then y, then z)."""
    rx, ry, rz = [float(v) for v in e]
    cx, sx, cy, sy, cz, sz = math.cos(rx), math.sin(rx), math.cos(ry), math.sin(ry), math.cos(rz), math.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


def _floats(s: Optional[str], n: Optional[int] = None, default=None):
    if s is None:
        return default
    v = [float(t) for t in s.split()]
    if n is not None and len(v) != n:
        raise ValueError(f"expected {n} floats, got {s!r}")
    return v


@dataclass
class Joint:
    name: str
    type: str  # 'free' | 'hinge'
    axis: np.ndarray  # unit, body frame
    pos: np.ndarray
    limited: bool
    range: tuple  # (lower, upper) radians
    armature: float
    damping: float
    stiffness: float


@dataclass
class Geom:
    type: str  # 'box' | 'cylinder' | 'capsule' | 'sphere'
    pos: np.ndarray  # body frame
    rot: np.ndarray  # 3x3 geom->body
    size: np.ndarray  # box: half extents(3); cylinder/capsule: (radius, half_height); sphere: (radius,)
    friction: float = 1.0


@dataclass
class Body:
    name: str
    parent: int  # -1 for root
    pos: np.ndarray  # in parent frame
    rot: np.ndarray  # 3x3 body->parent
    mass: float = 0.0
    com: np.ndarray = field(default_factory=lambda: np.zeros(3))
    inertia: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))  # about COM, body frame
    has_inertial: bool = False
    joints: List[Joint] = field(default_factory=list)
    geoms: List[Geom] = field(default_factory=list)


@dataclass
class RobotModel:
    name: str
    bodies: List[Body]
    actuators: Dict[str, tuple]  # joint name -> (ctrl_lo, ctrl_hi, gear)

    @property
    def num_bodies(self) -> int:
        return len(self.bodies)

    @property
    def dof_joints(self) -> List[Joint]:
        return [j for b in self.bodies for j in b.joints if j.type == "hinge"]

    @property
    def num_dofs(self) -> int:
        return len(self.dof_joints)

    def body_index(self, name: str) -> int:
        for i, b in enumerate(self.bodies):
            if b.name == name:
                return i
        return -1

    def total_mass(self) -> float:
        return float(sum(b.mass for b in self.bodies))


class _Defaults:
    """<default class=...> tree: attribute dicts per element tag, inherited from the parent class."""

    def __init__(self):
        self.classes: Dict[str, Dict[str, Dict[str, str]]] = {"__root__": {}}

    def parse(self, node, parent="__root__"):
        cls = node.get("class", parent if parent != "__root__" else "__root__")
        if node.get("class") is None and parent == "__root__":
            cls = "__root__"
        base = {k: dict(v) for k, v in self.classes.get(parent, {}).items()}
        self.classes[cls] = base
        for child in node:
            if child.tag == "default":
                continue
            base.setdefault(child.tag, {}).update(child.attrib)
        for child in node:
            if child.tag == "default":
                self.parse(child, cls)

    def resolve(self, tag: str, elem, childclass: Optional[str]) -> Dict[str, str]:
        cls = elem.get("class") or childclass or "__root__"
        out = dict(self.classes.get(cls, self.classes["__root__"]).get(tag, {}))
        out.update(elem.attrib)
        return out


def _geom_inertia(g: Geom, density: float):
    """mass, com(body frame), inertia about com (body frame) of a primitive at uniform density."""
    if g.type == "box":
        a, b, c = g.size
        m = density * 8 * a * b * c
        I = np.diag([m / 3 * (b * b + c * c), m / 3 * (a * a + c * c), m / 3 * (a * a + b * b)])
    elif g.type == "sphere":
        r = g.size[0]
        m = density * 4 / 3 * math.pi * r ** 3
        I = np.eye(3) * 0.4 * m * r * r
    elif g.type == "cylinder":
        r, h = g.size[0], g.size[1]
        m = density * math.pi * r * r * 2 * h
        I = np.diag([m * (3 * r * r + 4 * h * h) / 12, m * (3 * r * r + 4 * h * h) / 12, m * r * r / 2])
    elif g.type == "capsule":
        r, h = g.size[0], g.size[1]
        mc = density * math.pi * r * r * 2 * h
        ms = density * 4 / 3 * math.pi * r ** 3
        m = mc + ms
        izz = mc * r * r / 2 + ms * 0.4 * r * r
        ixx = mc * (3 * r * r + 4 * h * h) / 12 + ms * (0.4 * r * r + h * h + 0.75 * r * h)
        I = np.diag([ixx, ixx, izz])
    else:
        raise ValueError(g.type)
    return m, g.pos.copy(), g.rot @ I @ g.rot.T


def load_mjcf(path: str, *, balance_inertia: str = "off", geom_density: float = 1000.0,
              infer_missing_inertia: bool = False) -> RobotModel:
    """Parse an MJCF file.

    balance_inertia: 'off' (use <inertial> as given -- PhysX-like default) or 'mujoco'
        (replace diagonals that violate the triangle inequality by their mean, what
        <compiler balanceinertia="true"> does in MuJoCo).
    infer_missing_inertia: bodies without <inertial> get mass from their collision geoms at
        `geom_density` (Humanoid-style assets). Default False: the TOCABI reward hard-codes
        104.48 kg = sum of explicit <inertial> masses (dyros_dynamic_walk.py:917).
    """
    root = ET.parse(path).getroot()
    compiler = root.find("compiler")
    angle_deg = True
    if compiler is not None and compiler.get("angle", "degree") == "radian":
        angle_deg = False
    ang = math.pi / 180.0 if angle_deg else 1.0

    defaults = _Defaults()
    dnode = root.find("default")
    if dnode is not None:
        defaults.parse(dnode)

    bodies: List[Body] = []

    def parse_rot(attrs) -> np.ndarray:
        if "quat" in attrs:
            return quat_wxyz_to_mat(_floats(attrs["quat"], 4))
        if "euler" in attrs:
            return euler_xyz_to_mat([a * ang for a in _floats(attrs["euler"], 3)])
        return np.eye(3)

    def parse_body(node, parent_idx: int, childclass: Optional[str]):
        childclass = node.get("childclass", childclass)
        b = Body(name=node.get("name", f"body{len(bodies)}"), parent=parent_idx,
                 pos=np.array(_floats(node.get("pos"), 3, [0.0, 0.0, 0.0])), rot=parse_rot(node.attrib))
        idx = len(bodies)
        bodies.append(b)
        inert = node.find("inertial")
        if inert is not None:
            b.has_inertial = True
            b.mass = float(inert.get("mass"))
            b.com = np.array(_floats(inert.get("pos"), 3, [0.0, 0.0, 0.0]))
            if inert.get("fullinertia") is not None:
                xx, yy, zz, xy, xz, yz = _floats(inert.get("fullinertia"), 6)
                I = np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]])
            else:
                I = np.diag(_floats(inert.get("diaginertia"), 3, [0.0, 0.0, 0.0]))
            if inert.get("quat") is not None:
                Rq = quat_wxyz_to_mat(_floats(inert.get("quat"), 4))
                I = Rq @ I @ Rq.T
            if balance_inertia == "mujoco":
                w, V = np.linalg.eigh(I)
                if w[2] > w[0] + w[1] + 1e-12:
                    w = np.full(3, w.mean())
                    I = V @ np.diag(w) @ V.T
            b.inertia = I
        for jn in list(node.findall("joint")) + list(node.findall("freejoint")):
            a = defaults.resolve("joint", jn, childclass)
            jtype = "free" if jn.tag == "freejoint" else a.get("type", "hinge")
            axis = np.array(_floats(a.get("axis"), 3, [0.0, 0.0, 1.0]))
            nrm = np.linalg.norm(axis)
            axis = axis / nrm if nrm > 0 else axis
            rng = _floats(a.get("range"), 2, [0.0, 0.0])
            limited = a.get("limited", "auto")
            limited = (limited == "true") or (limited == "auto" and "range" in a)
            b.joints.append(Joint(
                name=a.get("name", f"joint{idx}"), type=jtype, axis=axis,
                pos=np.array(_floats(a.get("pos"), 3, [0.0, 0.0, 0.0])), limited=limited,
                range=(rng[0] * ang, rng[1] * ang) if jtype == "hinge" else (0.0, 0.0),
                armature=float(a.get("armature", 0.0)), damping=float(a.get("damping", 0.0)),
                stiffness=float(a.get("stiffness", 0.0))))
        for gn in node.findall("geom"):
            a = defaults.resolve("geom", gn, childclass)
            gtype = a.get("type", "sphere")
            if gtype not in ("box", "cylinder", "capsule", "sphere"):
                continue  # mesh/plane: visual only on this path
            if a.get("contype", "1") == "0" and a.get("conaffinity", "1") == "0":
                continue
            size = _floats(a.get("size"), None, [])
            pos = np.array(_floats(a.get("pos"), 3, [0.0, 0.0, 0.0]))
            rot = parse_rot(a)
            if "fromto" in a and gtype in ("cylinder", "capsule"):
                ft = np.array(_floats(a["fromto"], 6))
                p0, p1 = ft[:3], ft[3:]
                pos = 0.5 * (p0 + p1)
                d = p1 - p0
                hl = 0.5 * np.linalg.norm(d)
                z = d / (2 * hl)
                x = np.cross([0.0, 1.0, 0.0], z)
                if np.linalg.norm(x) < 1e-8:
                    x = np.cross([1.0, 0.0, 0.0], z)
                x /= np.linalg.norm(x)
                rot = np.stack([x, np.cross(z, x), z], axis=1)
                size = [size[0], hl]
            fr = _floats(a.get("friction"), None, [1.0])[0]
            b.geoms.append(Geom(type=gtype, pos=pos, rot=rot, size=np.array(size, dtype=np.float64), friction=fr))
        if not b.has_inertial and infer_missing_inertia and b.geoms:
            ms, cs, Is = zip(*[_geom_inertia(g, geom_density) for g in b.geoms])
            m = sum(ms)
            c = sum(mi * ci for mi, ci in zip(ms, cs)) / m
            I = np.zeros((3, 3))
            for mi, ci, Ii in zip(ms, cs, Is):
                d = ci - c
                I += Ii + mi * (d @ d * np.eye(3) - np.outer(d, d))
            b.mass, b.com, b.inertia = m, c, I
        for child in node.findall("body"):
            parse_body(child, idx, childclass)

    world = root.find("worldbody")
    tops = world.findall("body")
    if len(tops) != 1:
        raise ValueError("expected exactly one top-level <body> (one articulation per asset)")
    parse_body(tops[0], -1, None)

    actuators: Dict[str, tuple] = {}
    act = root.find("actuator")
    if act is not None:
        for m in act.findall("motor"):
            a = defaults.resolve("motor", m, None)
            rng = _floats(a.get("ctrlrange"), 2, [-math.inf, math.inf])
            gear = _floats(a.get("gear"), None, [1.0])[0]
            actuators[a["joint"]] = (rng[0], rng[1], gear)
    return RobotModel(name=root.get("model", "robot"), bodies=bodies, actuators=actuators)
