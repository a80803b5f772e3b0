import sys, torch
sys.path.insert(0, '.')
from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
env = DyrosDynamicWalk(default_cfg(4096), "cuda:0", use_cuda_graph=False)
g = torch.Generator(device="cuda:0"); g.manual_seed(1)
for i in range(8):
    env.step(torch.rand(4096, 13, device="cuda:0", generator=g) * 2 - 1)
torch.cuda.synchronize()
print("ok")
