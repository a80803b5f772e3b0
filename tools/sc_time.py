"""Back-to-back launches of k_self_collision alone on the state after a rollout: us per launch."""
import sys, torch
sys.path.insert(0, '.')
from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
N = 4096
env = DyrosDynamicWalk(default_cfg(N), "cuda:0")
g = torch.Generator(device="cuda:0"); g.manual_seed(1)
acts = [torch.rand(N, 13, device="cuda:0", generator=g) * 2 - 1 for _ in range(8)]
for warm in (4, 400):
    for i in range(warm):
        env.step(acts[i % 8])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        a.record()
        for i in range(200):
            env.core.self_collision()
        b.record(); torch.cuda.synchronize()
    print("after", warm, "steps: us/launch", a.elapsed_time(b) / 200 * 1e3)
