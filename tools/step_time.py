"""Back-to-back fused steps (CUDA graph, no L2 flush): ms per step. python tools/step_time.py [N] [steps]"""
import sys, time, torch
sys.path.insert(0, '.')
from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
pos = [x for x in sys.argv[1:] if not x.startswith("--")]
N = int(pos[0]) if len(pos) > 0 else 4096
K = int(pos[1]) if len(pos) > 1 else 300
cfg = default_cfg(N)
if "--no-self-collision" in sys.argv:
    cfg["env"]["selfCollision"] = False
env = DyrosDynamicWalk(cfg, "cuda:0")
g = torch.Generator(device="cuda:0"); g.manual_seed(1)
acts = [torch.rand(N, 13, device="cuda:0", generator=g) * 2 - 1 for _ in range(8)]
for i in range(400):
    env.step(acts[i % 8])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(K):
    env.step(acts[i % 8])
b.record(); torch.cuda.synchronize()
print("ms/step", a.elapsed_time(b) / K, "launches/step", env.core.step_launches())
