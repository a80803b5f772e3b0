"""Episode-length statistics under random actions with and without the self-collision pass:
   python tools/episode_lengths.py [steps] > profiles/r2_episode_lengths.md"""
import sys, torch
sys.path.insert(0, '.')
from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
N = 4096
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
print("# Episode lengths under random actions, 4096 envs x %d policy steps (seed 42), with / without k_self_collision\n" % STEPS)
print("| self-collision | episodes | mean length | median | p10 | p90 | max | resets with a non-foot self-contact > 1 N | resets with only ground / orientation causes |")
print("|---|---|---|---|---|---|---|---|---|")
for on in (True, False):
    cfg = default_cfg(N)
    cfg["env"]["selfCollision"] = on
    env = DyrosDynamicWalk(cfg, "cuda:0")
    g = torch.Generator(device="cuda:0"); g.manual_seed(42)
    acts = [torch.rand(N, 13, device="cuda:0", generator=g) * 2 - 1 for _ in range(16)]
    lens, self_hits, total = [], 0, 0
    feet = [b for b in range(38) if b in (8, 16)]
    nonfoot = torch.tensor([b for b in range(38) if b not in feet], device="cuda:0")
    prev_progress = env.progress_buf.clone()
    for t in range(STEPS):
        env.step(acts[t % 16])
        done = env.reset_buf != 0
        if t > 0 and bool(done.any()):
            # the env resets itself inside the step: its length is the progress counter it had before + 1
            lens.append((prev_progress[done] + 1).cpu())
            total += int(done.sum())
            if on:
                f = env.core.sim_t["self_contact_force"].view(N, 38, 3)[:, nonfoot].norm(dim=2).max(dim=1).values
                self_hits += int((done & (f > 1.0)).sum())
        prev_progress = env.progress_buf.clone()
    L = torch.cat(lens).float()
    q = lambda p: float(torch.quantile(L, p))
    print(f"| {'on' if on else 'off'} | {total} | {L.mean():.1f} | {q(0.5):.0f} | {q(0.1):.0f} | {q(0.9):.0f} | {int(L.max())} | "
          f"{self_hits if on else '-'} ({100.0 * self_hits / max(total, 1):.1f} %) | {total - self_hits} |")
    env.close()
print("\nRandom actions (uniform in [-1, 1] on 12 leg torques + the phase action) make the robot fall within ~0.1-0.3 s; the")
print("self-collision pass ends an episode earlier whenever a leg / arm link meets another before a non-foot body meets the ground.")
