"""isaacgym.gymtorch -> isaacgymdyros_b200.gymtorch (see the package docstring)."""
from isaacgymdyros_b200.gymtorch import unwrap_tensor, wrap_tensor  # noqa: F401
