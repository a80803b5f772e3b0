"""`gymtorch` facade (reference: python/isaacgym/gymtorch.py:61-106 and _bindings/src/gymtorch/gymtorch.cpp:33-158).

In the reference, `wrap_tensor` builds a non-owning torch view over memory owned by the simulator (no deleter,
gymtorch.cpp:90,121) and `unwrap_tensor` builds a descriptor over a torch tensor. Here the simulator's buffers ARE torch
allocations, so wrapping returns the very tensor the library writes into and lifetime is ordinary refcounting."""
import torch

from . import gymapi

_SUPPORTED = (torch.float32, torch.int32, torch.int64, torch.uint8, torch.int16)


def wrap_tensor(gym_tensor, offsets=None, counts=None):
    """gymtorch.py:61-94. Slicing (`offsets` / `counts`) returns a view, like the reference's strided wrap."""
    if gym_tensor is None or getattr(gym_tensor, "torch_tensor", None) is None or gym_tensor.torch_tensor.numel() == 0:
        print("*** Can't create empty tensor")  # gymtorch.cpp:40-45
        return None
    t = gym_tensor.torch_tensor
    if offsets is not None or counts is not None:
        offsets = offsets or [0] * t.dim()
        counts = counts or [s - o for s, o in zip(t.shape, offsets)]
        t = t[tuple(slice(o, o + c) for o, c in zip(offsets, counts))]
    return t


def unwrap_tensor(torch_tensor):
    """gymtorch.py:97-106: contiguous tensors of a supported dtype only."""
    if not torch_tensor.is_contiguous():
        raise Exception("Input tensor must be contiguous")  # gymtorch.py:98-99
    if torch_tensor.dtype not in _SUPPORTED:
        raise Exception("Unsupported Gym tensor dtype")  # gymtorch.py:85
    return gymapi.Tensor(torch_tensor, own_data=False)
