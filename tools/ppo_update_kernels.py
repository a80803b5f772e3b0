"""Kernel list of PPO minibatch updates (eager, no graph): run under
   ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file out.csv python tools/ppo_update_kernels.py"""
import sys, torch
sys.path.insert(0, '.')
from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
from isaacgymdyros_b200.ppo import PPOConfig, PPOTrainer
N = 4096
env = DyrosDynamicWalk(default_cfg(N), "cuda:0", use_cuda_graph=False)
tr = PPOTrainer(env, PPOConfig(horizon_length=8, minibatch_size=4096, use_cuda_graph=False))
tr.rollout(); tr._prepare()
print('packed' if tr.packed is not None else 'autograd')
for i in range(3):
    tr._minibatch(i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr._minibatch(3)
tr._minibatch(4)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(8):
    tr._minibatch(i)
b.record(); torch.cuda.synchronize()
print("eager ms/minibatch", a.elapsed_time(b) / 8)
