"""B200-native DyrosDynamicWalk env-step path. Import surface:
    from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
"""
__all__ = ["DyrosDynamicWalk", "default_cfg", "DyrosCore", "CoreConfig"]


def __getattr__(name):
    if name in ("DyrosDynamicWalk", "default_cfg"):
        from .tasks import dyros_dynamic_walk as m
        return getattr(m, name)
    if name in ("DyrosCore", "CoreConfig"):
        from . import core as m
        return getattr(m, name)
    raise AttributeError(name)
