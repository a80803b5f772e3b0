"""isaacgym.gymdeps: the reference preloads libcuda / PhysX / USD shared libraries here and refuses to be imported after
torch (gymdeps.py:20-60). libdyros_b200.so needs neither: it is loaded through ctypes on first use and shares torch's
CUDA context, so importing torch first is fine."""


def _import_deps():
    return None
