"""Minimal stand-in for gym.spaces.Box (the reference imports OpenAI gym, vec_task.py:33, which is not
installed here). If `gym` is importable its Box is used so rl_games sees the type it expects."""
import numpy as np

try:  # pragma: no cover - depends on the host environment
    from gym.spaces import Box  # type: ignore
except Exception:  # noqa: BLE001
    class Box:
        def __init__(self, low, high, dtype=np.float32):
            self.low = np.asarray(low, dtype=dtype)
            self.high = np.asarray(high, dtype=dtype)
            self.shape = self.low.shape
            self.dtype = np.dtype(dtype)

        def __repr__(self):
            return f"Box{self.shape}"
