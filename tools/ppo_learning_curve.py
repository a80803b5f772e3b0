"""On-device PPO for a few hundred epochs on one GPU (4096 envs x 128 steps per epoch, the reference's hyper-parameters):
episode statistics every 20 epochs, plus a finite-state check of the env.   python tools/ppo_learning_curve.py [epochs]"""
import sys, time, torch
sys.path.insert(0, '.')
from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
from isaacgymdyros_b200.ppo import PPOConfig, PPOTrainer
E = int(sys.argv[1]) if len(sys.argv) > 1 else 400
env = DyrosDynamicWalk(default_cfg(4096), "cuda:0", use_cuda_graph=False)
tr = PPOTrainer(env, PPOConfig())
print("# On-device PPO, one B200, 4096 envs x 128 steps per epoch (DyrosDynamicWalkPPO.yaml hyper-parameters, self-collision on)\n")
print("| epoch | frames | mean episode reward | mean episode length | episodes ended | actor loss | critic loss | kl | lr |")
print("|---|---|---|---|---|---|---|---|---|")
t0 = time.time()
for ep in range(1, E + 1):
    out = tr.train_epoch()
    if ep % 20 == 0 or ep == 1:
        print(f"| {ep} | {ep * 4096 * 128:,} | {out['mean_reward']:.1f} | {out['mean_length']:.1f} | {int(out['episodes'])} | "
              f"{out['a_loss']:.4f} | {out['c_loss']:.3f} | {out['kl']:.5f} | {out['lr']:.2e} |", flush=True)
torch.cuda.synchronize()
dt = time.time() - t0
finite = bool(torch.isfinite(env.core.sim_t["root_states"]).all() and torch.isfinite(env.core.sim_t["dof_state"]).all() and torch.isfinite(env.obs_buf).all())
print(f"\n{E} epochs = {E * 4096 * 128:,} env-steps in {dt:.1f} s wall ({E * 4096 * 128 / dt / 1e6:.2f} M env-steps/s including the host loop); "
      f"all env state finite at the end: {finite}.")
