"""isaacgym.gymapi -> isaacgymdyros_b200.gymapi (see the package docstring)."""
from isaacgymdyros_b200.gymapi import *  # noqa: F401,F403
from isaacgymdyros_b200.gymapi import acquire_gym  # noqa: F401
