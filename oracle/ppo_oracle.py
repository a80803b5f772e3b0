"""TEST INFRASTRUCTURE ONLY -- plain-torch restatement of the PPO pieces of the reference's rl_games fork that
isaacgymdyros_b200/ppo.py + csrc/ppo_kernels.cu implement on the device. Never imported by the product.

Follows learning/rl_games_custom/a2c_common_dyros.py (A2C), a2c_continuous_seperate.py (AG), models_dyros.py (MD) of the
reference, plus three functions of rl-games 1.1.4 (pinned by IsaacGymEnvs/setup.py:22, not vendored and not installed
here; restated from their published definitions): common_losses.actor_loss / critic_loss, torch_ext.policy_kl,
schedulers.LinearScheduler. Parity with rl_games itself is therefore unpinned; what the tests pin is product == this."""
import math

import torch


def neglogp(x, mean, logstd):
    """MD:60-63."""
    std = torch.exp(logstd)
    return 0.5 * (((x - mean) / std) ** 2).sum(dim=-1) + 0.5 * math.log(2.0 * math.pi) * x.size()[-1] + logstd.sum(dim=-1)


def discount_values(fdones, last_values, mb_fdones, mb_values, mb_rewards, gamma, tau):
    """A2C:485-500, time-major (H, N) tensors."""
    H = mb_rewards.shape[0]
    lastgaelam = 0
    mb_advs = torch.zeros_like(mb_rewards)
    for t in reversed(range(H)):
        if t == H - 1:
            nextnonterminal = 1.0 - fdones
            nextvalues = last_values
        else:
            nextnonterminal = 1.0 - mb_fdones[t + 1]
            nextvalues = mb_values[t + 1]
        delta = mb_rewards[t] + gamma * nextvalues * nextnonterminal - mb_values[t]
        mb_advs[t] = lastgaelam = delta + gamma * tau * nextnonterminal * lastgaelam
    return mb_advs


def actor_loss(old_neglogp, new_neglogp, advantage, e_clip):
    """rl_games common_losses.actor_loss (ppo=True) + the fork's clip fraction (AG:146)."""
    ratio = torch.exp(old_neglogp - new_neglogp)
    surr1 = advantage * ratio
    surr2 = advantage * torch.clamp(ratio, 1.0 - e_clip, 1.0 + e_clip)
    return torch.max(-surr1, -surr2), (torch.abs(ratio - 1.0) > e_clip).float().mean()


def critic_loss(values, returns):
    """rl_games common_losses.critic_loss with clip_value False (PPO:89)."""
    return (returns - values) ** 2


def policy_kl(p0_mu, p0_sigma, p1_mu, p1_sigma):
    """rl_games torch_ext.policy_kl, reduce=True."""
    c1 = torch.log(p1_sigma / p0_sigma + 1e-5)
    c2 = (p0_sigma ** 2 + (p1_mu - p0_mu) ** 2) / (2.0 * (p1_sigma ** 2 + 1e-5))
    return (c1 + c2 - 0.5).sum(dim=-1).mean()


def total_loss(mu, values, logstd, batch, e_clip, critic_coef):
    """AG:128-157 with entropy_coef = bounds_loss_coef = 0 (PPO:81,92). Returns loss and the logged scalars."""
    nl = neglogp(batch["actions"], mu, logstd)
    a, clip_frac = actor_loss(batch["old_neglogp"], nl, batch["advantages"], e_clip)
    c = critic_loss(values, batch["returns"])
    a_loss, c_loss = a.mean(), c.mean()
    sigma = torch.exp(logstd).expand_as(mu)
    kl = policy_kl(mu.detach(), sigma, batch["old_mu"], sigma)
    return a_loss + 0.5 * c_loss * critic_coef, a_loss.detach(), c_loss.detach(), kl, clip_frac


def linear_lr(step, lr0, lr_min, max_steps):
    """rl_games schedulers.LinearScheduler after `step` updates."""
    return lr_min + (lr0 - lr_min) * (max(0, max_steps - step) / max_steps)
