"""Golden vectors for the domain-randomisation tables, produced by the UNMODIFIED reference helper
`isaacgym/gymutil.py::apply_random_samples` (gymutil.py:584-619) with the parameters of DyrosDynamicWalk.yaml:78-115.
Run in the build container only:

    python tests/golden/make_dr_golden.py

The reference samples with numpy's global generator (gymutil.py:549-567). Here `np.random.uniform` is replaced, for the
duration of the calls, by a recorder that draws u on the 24-bit grid the CUDA kernels use and returns lo + (hi - lo) * u
(numpy's own formula), so the same u can be injected into the kernels (DyrosNoiseInjection.dr_u). Each env's property
array is randomised TWICE from the same originals: the second result must not compound the first (gymutil.py:602-605
use `og_prop`, vec_task.py:686-704).
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DYROS_REFERENCE_ROOT", "/root/reference")
ARMATURE = [0.614, 0.862, 1.09, 1.09, 1.09, 0.360, 0.614, 0.862, 1.09, 1.09, 1.09, 0.360,
            0.078, 0.078, 0.078, 0.18, 0.18, 0.18, 0.18, 0.0032, 0.0032, 0.0032, 0.0032, 0.0032, 0.0032,
            0.18, 0.18, 0.18, 0.18, 0.0032, 0.0032, 0.0032, 0.0032]          # dyros_dynamic_walk.py:366-371
DOF_DTYPE = np.dtype([("hasLimits", "?"), ("lower", "f4"), ("upper", "f4"), ("driveMode", "i4"), ("velocity", "f4"),
                      ("effort", "f4"), ("stiffness", "f4"), ("damping", "f4"), ("friction", "f4"), ("armature", "f4")])
# DyrosDynamicWalk.yaml:81-115
P_MASS = {"range": [0.8, 1.2], "operation": "scaling", "distribution": "uniform", "setup_only": True,
          "schedule": "constant", "schedule_steps": 0}
P_DAMP = {"range": [0.0, 2.9], "operation": "additive", "distribution": "uniform", "schedule": "constant", "schedule_steps": 0}
P_ARM = {"range": [0.8, 1.2], "operation": "scaling", "distribution": "uniform", "schedule": "constant", "schedule_steps": 0}


def load_gymutil():
    pkg = types.ModuleType("isaacgym")
    pkg.__path__ = []
    gymapi = types.ModuleType("isaacgym.gymapi")

    class SimParams:  # only the isinstance test of gymutil.py:586 needs it
        pass

    gymapi.SimParams = SimParams
    saved = {k: sys.modules.get(k) for k in ("isaacgym", "isaacgym.gymapi")}
    sys.modules["isaacgym"], sys.modules["isaacgym.gymapi"] = pkg, gymapi
    try:
        spec = importlib.util.spec_from_file_location("isaacgym.gymutil", os.path.join(REF, "python", "isaacgym", "gymutil.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


class UniformRecorder:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.log = []

    def __enter__(self):
        self._orig = np.random.uniform

        def uniform(lo, hi, shape):
            u = self.rng.integers(0, 1 << 24, size=shape).astype(np.float64) / float(1 << 24)
            self.log.append(np.asarray(u).copy())
            return lo + (hi - lo) * u

        np.random.uniform = uniform
        return self

    def __exit__(self, *exc):
        np.random.uniform = self._orig


def main():
    GU = load_gymutil()
    N, nd, nb = 12, 33, 38
    z = np.load(os.path.join(HERE, "..", "..", "isaacgymdyros_b200", "assets", "tocabi_tables.npz"))
    mass0 = z["body_inertia"][:, 0].astype(np.float32)
    og = np.zeros(nd, dtype=DOF_DTYPE)
    og["damping"] = 0.1                                  # dyros_dynamic_walk.py:365
    og["armature"] = np.array(ARMATURE, np.float32)      # :366-371
    og["velocity"] = 4.03
    out = {"meta_N": N}
    prev = {}
    with UniformRecorder(2024) as rec:
        for rnd in range(2):
            damp, arm, u_d, u_a = [], [], [], []
            for e in range(N):
                prop = og.copy() if rnd == 0 else prev[e]
                n0 = len(rec.log)
                # vec_task.py:698-707: for attr in prop_attrs (yaml order: damping, armature), randomization_ct = last_step
                GU.apply_random_samples(prop, og, "damping", P_DAMP, 5 + rnd)
                GU.apply_random_samples(prop, og, "armature", P_ARM, 5 + rnd)
                assert len(rec.log) == n0 + 2
                u_d.append(rec.log[n0]); u_a.append(rec.log[n0 + 1])
                damp.append(prop["damping"].copy()); arm.append(prop["armature"].copy())
                prev[e] = prop
            out[f"r{rnd}/u_damping"], out[f"r{rnd}/u_armature"] = np.array(u_d, np.float32), np.array(u_a, np.float32)
            out[f"r{rnd}/damping"], out[f"r{rnd}/armature"] = np.array(damp, np.float32), np.array(arm, np.float32)
        # rigid_body_properties.mass: a list of objects per env, one scalar draw per body (gymutil.py:607-619)
        u_m, mass = [], []
        for e in range(N):
            props = [types.SimpleNamespace(mass=float(m)) for m in mass0]
            ogp = [{"mass": float(m)} for m in mass0]
            n0 = len(rec.log)
            for p, o in zip(props, ogp):
                GU.apply_random_samples(p, o, "mass", P_MASS, 0)
            u_m.append(np.concatenate([np.ravel(x) for x in rec.log[n0:]]))
            mass.append([float(np.ravel(p.mass)[0]) for p in props])
        out["u_mass"], out["mass"], out["mass0"] = np.array(u_m, np.float32), np.array(mass, np.float32), mass0
    path = os.path.join(HERE, "dr_samples.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
