import sys, torch
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sys.path.insert(0, '.')
from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
env = DyrosDynamicWalk(default_cfg(4096), "cuda:0", use_cuda_graph=False)
g = torch.Generator(device="cuda:0"); g.manual_seed(1)
for i in range(STEPS):
    env.step(torch.rand(4096, 13, device="cuda:0", generator=g) * 2 - 1)
torch.cuda.synchronize()
print("ok")
