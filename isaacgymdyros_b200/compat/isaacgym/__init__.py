"""`isaacgym` as the reference's Python code imports it, backed by isaacgymdyros_b200.

Put this directory's parent (isaacgymdyros_b200/compat) at the FRONT of sys.path / PYTHONPATH. The modules that bind the
closed native runtime are replaced here:

    isaacgym.gymapi    <- isaacgymdyros_b200.gymapi   (reference: python/isaacgym/gymapi.py:32-101 loads gym_3x.so)
    isaacgym.gymtorch  <- isaacgymdyros_b200.gymtorch (reference: gymtorch.py JIT-builds _bindings/src/gymtorch)
    isaacgym.gymdeps   <- nothing to preload          (reference: gymdeps.py:20-60 preloads PhysX / USD libraries)

Everything else of the Isaac Gym Python package is plain Python with no native dependency (torch_utils.py, gymutil.py,
terrain_utils.py) and keeps coming from the user's own Isaac Gym checkout: set DYROS_ISAACGYM_PY to its
`python/isaacgym` directory and those submodules resolve there (this package's __path__ is extended with it, after this
directory, so the three modules above win)."""
import os

_user = os.environ.get("DYROS_ISAACGYM_PY")
if _user:
    if not os.path.isfile(os.path.join(_user, "torch_utils.py")):
        raise ImportError(f"DYROS_ISAACGYM_PY={_user!r} does not look like Isaac Gym's python/isaacgym directory")
    __path__.append(_user)
