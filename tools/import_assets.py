"""One-off asset import (run in the build container where /root/reference is mounted).

Converts the reference's data files for this path into this repo's own binary formats so that the
GPU box (which has no /root/reference) can run:
  assets/mjcf/dyros_tocabi/xml/dyros_tocabi.xml  -> assets/tocabi_tables.npz  (flat model tables, our layout)
  assets/mjcf/nv_humanoid.xml                    -> assets/humanoid_tables.npz (BASELINE configs[2] generality check)
  assets/DeepMimic/processed_data_tocabi_walk.txt -> assets/mocap_walk.npy     (3600x36 float32, as the
        reference casts it: dyros_dynamic_walk.py:112-113)
  assets/Data/obs_{mean,variance}_fixed.txt       -> assets/obs_norm.npy       (2x37 float32, :139-142)
The MJCF loader itself (model/mjcf.py) works on any user-supplied path at run time; the .npz is the
fallback used when no IsaacGymEnvs asset tree is present.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from isaacgymdyros_b200.model.mjcf import load_mjcf  # noqa: E402
from isaacgymdyros_b200.model.tables import build_tables  # noqa: E402

REF = os.environ.get("DYROS_REF_ASSETS", "/root/reference/python/IsaacGymEnvs/assets")
OUT = os.path.join(os.path.dirname(__file__), "..", "isaacgymdyros_b200", "assets")


def main():
    os.makedirs(OUT, exist_ok=True)
    m = load_mjcf(os.path.join(REF, "mjcf/dyros_tocabi/xml/dyros_tocabi.xml"))
    t = build_tables(m, solver_bodies=["L_Foot_Link", "R_Foot_Link"])
    t.save(os.path.join(OUT, "tocabi_tables.npz"))
    hm = load_mjcf(os.path.join(REF, "mjcf/nv_humanoid.xml"), infer_missing_inertia=True)
    ht = build_tables(hm, solver_bodies=["right_foot", "left_foot"], vel_limit=1.0e3)
    ht.save(os.path.join(OUT, "humanoid_tables.npz"))
    print("humanoid: bodies", ht.num_bodies, "links", ht.num_links, "dofs", ht.num_dofs, "mass", ht.total_mass())
    mocap = np.genfromtxt(os.path.join(REF, "DeepMimic/processed_data_tocabi_walk.txt"), encoding="ascii")
    assert mocap.shape == (3600, 36), mocap.shape
    np.save(os.path.join(OUT, "mocap_walk.npy"), mocap.astype(np.float32))
    mean = np.genfromtxt(os.path.join(REF, "Data/obs_mean_fixed.txt"), encoding="ascii")
    var = np.genfromtxt(os.path.join(REF, "Data/obs_variance_fixed.txt"), encoding="ascii")
    np.save(os.path.join(OUT, "obs_norm.npy"), np.stack([mean, var]).astype(np.float32))
    print("bodies", t.num_bodies, "links", t.num_links, "dofs", t.num_dofs, "mass", t.total_mass(),
          "mocap", mocap.shape, "obs_norm", mean.shape)


if __name__ == "__main__":
    main()
