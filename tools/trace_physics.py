"""Phase timeline of the physics kernel (CTA 0) from the clock64() marks of dyros_task_physics_trace.
    python tools/trace_physics.py  -> table of cycles per phase and role (run on a GPU box)"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg

NAMES = ["pass1", "wait-children-pass1", "pass2", "base solve", "foot G-walk", "pass3", "foot V/Om/rows", "sweeps",
         "impulse", "base response", "down pass", "base integrate"]
env = DyrosDynamicWalk(default_cfg(4096), "cuda:0", use_cuda_graph=False)
g = torch.Generator(device="cuda:0"); g.manual_seed(1)
for _ in range(30):
    env.step(torch.rand(4096, 13, device="cuda:0", generator=g) * 2 - 1)
acts = torch.rand(4096, 13, device="cuda:0", generator=g) * 2 - 1
fused = "--fused" in sys.argv  # prologue inside the physics launch (what the fused step does) instead of its own kernel
if not fused:
    env.core.prologue(acts)
if "--cold" in sys.argv:  # evict everything from L2 first (what bench.py does between timed steps)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
    flush.zero_()
    torch.cuda.synchronize()
tr = (env.core.prologue_physics(acts, trace=True) if fused else env.core.task_physics_trace()).cpu()
torch.cuda.synchronize()
k0 = int(tr[0, 0, 24])
print(f"kernel entry 0, tables staged / previous kernel done {int(tr[0, 0, 25]) - k0}, sub-step starts "
      f"{[int(tr[s, :, 13].min()) - k0 for s in range(tr.shape[0])]}, all outputs written {int(tr[0, 0, 26]) - k0}")
for s in range(tr.shape[0]):
    t0 = int(tr[s, :, 13].min())
    print(f"sub-step {s}: roles start at {[int(tr[s, r, 0]) - t0 for r in range(4)]}, end {[int(tr[s, r, 15]) - t0 for r in range(4)]}")
    io = [int(tr[s, 0, i]) - t0 for i in range(18, 23)]
    print(f"  I/O group: start {io[0]}, push+zero staged (F_IO_PRE) {io[1]}, prologue done {io[2]}, torque stage done {io[3]}, F_IO_TAU {io[4]}")
    print(f"{'phase':24s}" + "".join(f"   role{r}: start   dur" for r in range(4)))
    print("pass1 split (E loop | propagation | forces): " + "  ".join(
        f"role{r}: {int(tr[s, r, 16]) - int(tr[s, r, 0])} | {int(tr[s, r, 17]) - int(tr[s, r, 16])} | {int(tr[s, r, 1]) - int(tr[s, r, 17])}" for r in range(4)))
    for i, n in enumerate(NAMES):
        print(f"{n:24s}" + "".join(f"   {int(tr[s, r, i]) - t0:12d} {int(tr[s, r, i + 1]) - int(tr[s, r, i]):6d}" for r in range(4)))
