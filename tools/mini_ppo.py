"""A tiny PPO epoch (64 envs x 4 steps, packed bf16 path, no CUDA graphs) for compute-sanitizer / ncu."""
import sys, torch
sys.path.insert(0, '.')
from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
from isaacgymdyros_b200.ppo import PPOConfig, PPOTrainer
env = DyrosDynamicWalk(default_cfg(64), "cuda:0", use_cuda_graph=False)
tr = PPOTrainer(env, PPOConfig(horizon_length=4, minibatch_size=128, mini_epochs=2, use_cuda_graph=False))
print(tr.train_epoch())
torch.cuda.synchronize()
print("ok")
