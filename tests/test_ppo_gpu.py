"""On-device PPO (isaacgymdyros_b200/ppo.py + csrc/ppo_kernels.cu) against the plain-torch restatement of
the reference's rl_games fork (oracle/ppo_oracle.py): rollout bookkeeping, GAE, loss gradients, the two Adam optimisers
with the actor-only gradient clip and the per-minibatch linear schedule, and whole minibatch updates (fp32 path: tight
tolerances; bf16 path: runs and stays close)."""
import math

import numpy as np
import pytest
import torch

from oracle import ppo_oracle as PO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_trainer(N=64, H=16, mb=256, **kw):
    from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
    from isaacgymdyros_b200.ppo import PPOConfig, PPOTrainer
    env = DyrosDynamicWalk(default_cfg(N), DEV, use_cuda_graph=False)
    cfg = PPOConfig(horizon_length=H, minibatch_size=mb, **kw)
    return env, PPOTrainer(env, cfg)


def test_parameter_counts_and_init():
    env, tr = make_trainer(N=8, H=4, mb=32)
    assert tr.net.n_actor == 194061 and tr.net.n - tr.net.n_actor == 190977  # 385,038 fp32 = 1.54 MB
    w, b = tr.net.layers["actor_mlp.0"]
    assert w.shape == (256, 487) and float(b.abs().max()) == 0.0
    wtw = (w @ w.t()) / 0.01 ** 2                                             # orthogonal rows, gain 0.01 (PPO:36-38)
    assert torch.allclose(wtw, torch.eye(256, device=DEV), atol=1e-3)
    assert torch.allclose(tr.net.logstd, torch.full((13,), -2.302585, device=DEV))
    tr.net.update_action_noise(0.75)
    assert torch.allclose(tr.net.logstd, torch.full((13,), 0.5 * -2.302585 + 0.5 * -2.9957, device=DEV))
    tr.net.update_action_noise(0.4)
    assert torch.allclose(tr.net.logstd, torch.full((13,), -2.9957, device=DEV))
    env.close()


def test_rollout_records_and_gae_match_oracle():
    N, H = 64, 16
    env, tr = make_trainer(N=N, H=H, mb=256, mixed_precision="fp32", use_cuda_graph=False)
    g = torch.Generator(device=DEV); g.manual_seed(0)
    z = torch.randn(H, N, 13, device=DEV, generator=g)
    tr.inject_normal = z.contiguous()
    # time-outs (value bootstrap, A2C:656-661). In the reference's flow an episode resets one step BEFORE its time-out
    # flag could rise (VT:325 tests the pre-increment counter, T:594 the incremented one), so the flag only fires from a
    # counter that already sits at the limit
    env.progress_buf[::7] = 7999
    obs_seen, done_seen, rew_seen, to_seen, val_seen, mu_seen = [], [], [], [], [], []
    for n in range(H):
        obs_seen.append(env.obs_buf.clone()); done_seen.append(env.reset_buf.clone())
        with torch.no_grad():
            mu, v = tr.net.forward(env.obs_buf)
        mu_seen.append(mu.clone()); val_seen.append(v.clone())
        tr._rollout_step()
        rew_seen.append(env.rew_buf.clone()); to_seen.append(env.timeout_buf.clone())
    torch.cuda.synchronize()
    b = tr.buf
    tm = lambda t: t.transpose(0, 1)                       # (N,H,.) -> (H,N,.)
    logstd = tr.net.logstd
    mu_s, val_s = torch.stack(mu_seen), torch.stack(val_seen)
    act = mu_s + torch.exp(logstd) * z
    assert torch.equal(tm(b["obs"]), torch.stack(obs_seen))
    assert torch.equal(tm(b["dones"]), (torch.stack(done_seen) != 0).float())
    assert torch.allclose(tm(b["actions"]), act, rtol=1e-6, atol=1e-7)
    assert torch.allclose(tm(b["neglogp"]), PO.neglogp(act, mu_s, logstd), rtol=1e-5, atol=1e-5)
    shaped = torch.stack(rew_seen) + 0.99 * val_s * (torch.stack(to_seen) != 0).float()
    assert torch.allclose(tm(b["rewards"]), shaped, rtol=1e-6, atol=1e-7)
    assert int((torch.stack(to_seen) != 0).sum()) > 0
    assert int(b["step"].item()) == 0 and int(b["global_step"].item()) == H
    # GAE + returns
    with torch.no_grad():
        _, last_v = tr.net.forward(env.obs_buf)
    import ctypes as C
    from isaacgymdyros_b200 import native
    native.check(tr.lib.dyros_ppo_gae(C.byref(tr.pb), tr._p(last_v.contiguous()), tr._p(env.reset_buf), tr._stream), "gae")
    torch.cuda.synchronize()
    want = PO.discount_values((env.reset_buf != 0).float(), last_v, tm(b["dones"]), tm(b["values"]), tm(b["rewards"]), 0.99, 0.95)
    assert torch.allclose(tm(b["advantages"]), want, rtol=1e-5, atol=1e-6)
    assert torch.allclose(b["returns"], b["advantages"] + b["values"], rtol=1e-6, atol=1e-7)
    env.close()


def fill_synthetic(tr, seed=1):
    g = torch.Generator(device=DEV); g.manual_seed(seed)
    r = lambda *s: torch.randn(*s, device=DEV, generator=g)
    b, N, H = tr.buf, tr.N, tr.H
    b["obs"].copy_(r(N, H, 487))
    b["mus"].copy_(0.3 * r(N, H, 13))
    b["actions"].copy_(b["mus"] + 0.12 * r(N, H, 13))
    b["neglogp"].copy_(PO.neglogp(b["actions"], b["mus"], tr.net.logstd) + 0.3 * r(N, H))
    b["values"].copy_(r(N, H)); b["returns"].copy_(b["values"] + r(N, H)); b["advantages"].copy_(b["returns"] - b["values"])
    tr._prepare()


def oracle_batch(tr, r0, mb):
    b, NH = tr.buf, tr.N * tr.H
    f = lambda k, *s: b[k].reshape(NH, *s)[r0:r0 + mb].clone()
    return {"actions": f("actions", 13), "old_neglogp": f("neglogp"), "advantages": tr.adv_norm.reshape(NH)[r0:r0 + mb].clone(),
            "returns": f("returns"), "old_mu": f("mus", 13)}


def test_loss_gradients_match_autograd_of_the_oracle():
    import ctypes as C
    from isaacgymdyros_b200 import native
    env, tr = make_trainer(N=64, H=16, mb=256, mixed_precision="fp32")
    fill_synthetic(tr)
    mb, r0 = 256, 512
    g = torch.Generator(device=DEV); g.manual_seed(3)
    mu = (tr.buf["mus"].reshape(-1, 13)[r0:r0 + mb] + 0.05 * torch.randn(mb, 13, device=DEV, generator=g)).requires_grad_(True)
    v = torch.randn(mb, device=DEV, generator=g).requires_grad_(True)
    batch = oracle_batch(tr, r0, mb)
    loss, a_loss, c_loss, kl, cf = PO.total_loss(mu, v, tr.net.logstd, batch, 0.2, 0.5)
    loss.backward()
    tr.stats.zero_()
    native.check(tr.lib.dyros_ppo_loss_grad(C.byref(tr.pb), r0, mb, tr._p(mu.detach().contiguous()), tr._p(v.detach().contiguous()),
                                            tr._p(tr.net.logstd), tr._p(tr.adv_norm), tr._p(tr.mb_dmu), tr._p(tr.mb_dv),
                                            tr._p(tr.stats), tr._stream), "loss_grad")
    torch.cuda.synchronize()
    assert torch.allclose(tr.mb_dmu, mu.grad, rtol=2e-4, atol=1e-8)
    assert torch.allclose(tr.mb_dv, v.grad, rtol=1e-5, atol=1e-9)
    got = tr.stats.tolist()
    for x, y in zip(got, (a_loss.item(), c_loss.item(), kl.item(), cf.item())):
        assert x == pytest.approx(y, rel=2e-4, abs=1e-6)
    assert 0.02 < cf.item() < 0.98  # both branches of the clipped surrogate are exercised
    env.close()


def test_adam_pair_with_actor_clip_and_linear_schedule_matches_torch():
    import ctypes as C
    from isaacgymdyros_b200 import native
    env, tr = make_trainer(N=8, H=4, mb=32, mixed_precision="fp32")
    n, na = tr.net.n, tr.net.n_actor
    ref = tr.net.flat.clone().requires_grad_(True)
    pa, pc = ref.detach()[:na].clone().requires_grad_(True), ref.detach()[na:].clone().requires_grad_(True)
    oa = torch.optim.Adam([pa], lr=1e-5, eps=1e-8)
    oc = torch.optim.Adam([pc], lr=5e-4, eps=1e-8)
    g = torch.Generator(device=DEV); g.manual_seed(5)
    world = 2
    for step in range(1, 6):
        grads = torch.randn(n, device=DEV, generator=g) * (10.0 if step % 2 else 1e-3)   # clipped and unclipped steps
        tr.net.grad.copy_(grads * world)                                                  # what an all-reduce SUM over 2 ranks leaves
        native.check(tr.lib.dyros_ppo_adam(tr._p(tr.net.flat), tr._p(tr.net.grad), tr._p(tr.net.exp_avg), tr._p(tr.net.exp_avg_sq),
                                           na, n, 1.0 / world, 0.5, tr._p(tr.norm2), tr._p(tr.lr), tr._p(tr.opt_step), 0.9, 0.999,
                                           1e-8, 1e-5, 3e-6, 5000, tr._stream), "adam")
        pa.grad, pc.grad = grads[:na].clone(), grads[na:].clone()
        torch.nn.utils.clip_grad_norm_([pa], 0.5)                                         # AG:186: the actor's parameters only
        oa.step(); oc.step()
        lr = PO.linear_lr(step, 1e-5, 3e-6, 5000)                                         # A2C:888-892, per minibatch
        for gp in oa.param_groups:
            gp["lr"] = lr
        torch.cuda.synchronize()
        assert tr.lr[0].item() == pytest.approx(lr, rel=1e-6) and tr.lr[1].item() == pytest.approx(5e-4)
        assert torch.allclose(tr.net.flat[:na], pa.detach(), rtol=1e-5, atol=1e-8)
        assert torch.allclose(tr.net.flat[na:], pc.detach(), rtol=1e-5, atol=1e-8)
    assert int(tr.opt_step.item()) == 5 and float(tr.norm2.item()) == 0.0
    env.close()


@pytest.mark.parametrize("graph", [False, True])
def test_minibatch_updates_track_the_oracle_fp32(graph):
    """Two mini-epochs over 4 minibatches through PPOTrainer.update (fp32 networks) against torch autograd + torch Adam on
    the same synthetic rollout: parameters, logged losses and the refreshed old-mu (dataset.update_mu_sigma, A2C:884)."""
    env, tr = make_trainer(N=64, H=16, mb=256, mixed_precision="fp32", mini_epochs=2, use_cuda_graph=graph)
    fill_synthetic(tr, seed=2)
    na = tr.net.n_actor
    flat0 = tr.net.flat.clone()
    mus0 = tr.buf["mus"].clone()
    # ---- oracle
    import copy
    from isaacgymdyros_b200.ppo import FlatActorCritic
    onet = FlatActorCritic(DEV, tr.cfg)
    onet.flat.copy_(flat0)
    params = [p for wb in onet.layers.values() for p in wb]
    actor_p = [p for k, wb in onet.layers.items() if k.startswith("actor") or k == "mu" for p in wb]
    critic_p = [p for k, wb in onet.layers.items() if k.startswith("critic") or k == "value" for p in wb]
    oa, oc = torch.optim.Adam(actor_p, lr=1e-5, eps=1e-8), torch.optim.Adam(critic_p, lr=5e-4, eps=1e-8)
    old_mu = mus0.reshape(-1, 13).clone()
    logs, step = [], 0
    for ep in range(2):
        for i in range(tr.num_minibatches):
            r0, mb = i * 256, 256
            batch = oracle_batch(tr, r0, mb)
            batch["old_mu"] = old_mu[r0:r0 + mb].clone()
            mu, v = onet.forward(tr.buf["obs"].reshape(-1, 487)[r0:r0 + mb])
            loss, a_loss, c_loss, kl, cf = PO.total_loss(mu, v, onet.logstd, batch, 0.2, 0.5)
            for p in params:
                p.grad = None
            loss.backward()
            torch.nn.utils.clip_grad_norm_(actor_p, 0.5)
            oa.step(); oc.step()
            step += 1
            for gp in oa.param_groups:
                gp["lr"] = PO.linear_lr(step, 1e-5, 3e-6, 5000)
            old_mu[r0:r0 + mb] = mu.detach()
            logs.append((a_loss.item(), c_loss.item(), kl.item(), cf.item()))
    # ---- product
    tr.update()
    torch.cuda.synchronize()
    assert torch.allclose(tr.net.flat, onet.flat, rtol=1e-4, atol=2e-7)
    assert float((tr.net.flat - flat0).abs().max()) > 1e-5          # the update moved the parameters at all
    want = np.mean(np.array(logs), axis=0)
    got = (tr.stats / (2 * tr.num_minibatches)).tolist()
    for x, y in zip(got, want):
        assert x == pytest.approx(y, rel=1e-3, abs=1e-6)
    assert torch.allclose(tr.buf["mus"].reshape(-1, 13), old_mu, rtol=1e-4, atol=1e-6)
    env.close()


def test_training_epochs_run_on_device_bf16():
    env, tr = make_trainer(N=256, H=16, mb=1024)
    env.progress_buf[::5] = 7990                                     # some episodes end inside the first rollout
    before = tr.net.flat.clone()
    episodes = 0
    for ep in range(3):
        out = tr.train_epoch()
        assert all(math.isfinite(v) for v in out.values()), out
        assert out["frames"] == 256 * 16
        episodes += out["episodes"]
        assert out["episodes"] == 0 or 0 < out["mean_length"] < 8000
    assert float((tr.net.flat - before).abs().max()) > 0
    assert episodes > 0
    assert out["lr"] < 1e-5                                          # the per-minibatch linear schedule moved
    env.close()


def test_packed_bf16_networks_match_the_fp32_networks():
    """The packed bf16 path (PackedNets + csrc 'packed bf16 path': batch-2 GEMMs, hand-written backward) against torch
    autograd on the fp32 master parameters, same minibatch: head outputs, loss gradients of every parameter, logged
    losses. Both hold identical (bf16-representable) weights, so what differs is the bf16 rounding of the activations:
    0.4 % per value, and ReLU masks that flip where a pre-activation is within that of zero (measured: 0.1 % of the
    critic's hidden units x rows, which alone costs 1e-3 of cosine); hence cosine > 0.997 and max error < 0.2 x the
    largest entry for the gradients, 2 % of the largest entry for the outputs."""
    import ctypes as C
    from isaacgymdyros_b200 import native
    from isaacgymdyros_b200.ppo import FlatActorCritic
    env, tr = make_trainer(N=64, H=16, mb=256)                       # mixed_precision defaults to bf16 -> packed
    assert tr.packed is not None
    pk = tr.packed
    g = torch.Generator(device=DEV); g.manual_seed(9)
    r = lambda *s: torch.randn(*s, device=DEV, generator=g)
    # larger-than-init weights so that every layer matters, non-zero biases
    # (rounded to bf16 so that both paths hold identical weights: what is compared is the arithmetic, not the cast)
    tr.net.flat.copy_((tr.net.flat * 30 + 0.02 * r(tr.net.n)).to(torch.bfloat16).float())
    pk.pack()
    N, H, mb, r0 = tr.N, tr.H, 256, 512
    obs = r(N * H, 487)
    tr.x_roll[:, :487] = obs.to(torch.bfloat16); tr.x_roll[:, 487] = 0
    b = tr.buf
    b["mus"].copy_(0.3 * r(N, H, 13)); b["actions"].copy_(b["mus"] + 0.12 * r(N, H, 13))
    b["neglogp"].copy_(PO.neglogp(b["actions"], b["mus"], tr.net.logstd) + 0.3 * r(N, H))
    b["values"].copy_(r(N, H)); b["returns"].copy_(b["values"] + r(N, H)); b["advantages"].copy_(b["returns"] - b["values"])
    tr._prepare()
    batch = oracle_batch(tr, r0, mb)
    # ---- fp32 autograd on the same (bf16-rounded) inputs
    import dataclasses
    onet = FlatActorCritic(DEV, dataclasses.replace(tr.cfg, mixed_precision="fp32"))
    onet.flat.copy_(tr.net.flat)
    x32 = tr.x_roll[r0:r0 + mb, :487].float()
    mu, v = onet.forward(x32)
    loss, a_loss, c_loss, kl, cf = PO.total_loss(mu, v, onet.logstd, batch, 0.2, 0.5)
    onet.grad.zero_()
    loss.backward()
    # ---- packed
    x = tr.x_roll[r0:r0 + mb]
    out = pk.forward(x)
    mu_p, v_p = pk.mu_value(out)
    scale = lambda t: float(t.abs().max())
    assert float((mu_p - mu).abs().max()) < 2e-2 * scale(mu) and float((v_p - v).abs().max()) < 2e-2 * scale(v)
    tr.stats.zero_()
    native.check(tr.lib.dyros_ppo_loss_grad_packed(C.byref(tr.pb), r0, mb, tr._p(out), tr._p(pk.bh), tr._p(tr.net.logstd), tr._p(tr.adv_norm),
                                                   tr._p(pk._buffers(mb)["dout"]), tr._p(pk.gbh), tr._p(tr.stats), tr._stream), "loss_grad_packed")
    pk.backward(x)
    pk.unpack_grads()
    torch.cuda.synchronize()
    assert float(pk.gb0.abs().max()) == 0.0 and float(pk.gbh.abs().max()) == 0.0      # accumulators are left zeroed
    assert torch.allclose(b["mus"].reshape(-1, 13)[r0:r0 + mb], mu_p, atol=1e-6)      # dataset.update_mu_sigma
    for name, (w, bias) in onet.layers.items():
        gw, gb = tr.net.layers[name][0].grad, tr.net.layers[name][1].grad
        for got, want, what in ((gw, w.grad, "weight"), (gb, bias.grad, "bias")):
            err = float((got - want).abs().max())
            cos = float((got * want).sum() / (got.norm() * want.norm() + 1e-30))
            assert err < 0.2 * scale(want) + 1e-12, (name, what, err, scale(want))
            assert cos > 0.997, (name, what, cos)
    for x_, y_ in zip(tr.stats.tolist(), (a_loss.item(), c_loss.item(), kl.item(), cf.item())):
        assert x_ == pytest.approx(y_, rel=5e-2, abs=2e-3)
    # ---- the optimiser on the packed path (norm folded into the unpack, pack folded into Adam) against dyros_ppo_adam
    grad = tr.net.grad.clone()
    ref = {k: t.clone() for k, t in (("p", tr.net.flat), ("m", tr.net.exp_avg), ("v", tr.net.exp_avg_sq))}
    lr2, st2, n2 = tr.lr.clone(), tr.opt_step.clone(), torch.zeros(1, device=DEV)
    grad.mul_(100.0)                                                 # large enough for the actor clip (0.5) to bite
    native.check(tr.lib.dyros_ppo_adam(tr._p(ref["p"]), tr._p(grad), tr._p(ref["m"]), tr._p(ref["v"]), tr.net.n_actor, tr.net.n, 0.5, 0.5,
                                       tr._p(n2), tr._p(lr2), tr._p(st2), 0.9, 0.999, 1e-8, 1e-5, 3e-6, 5000, tr._stream), "adam")
    tr.net.grad.copy_(grad)
    tr.norm2.copy_((grad[:tr.net.n_actor] ** 2).sum().reshape(1))    # what dyros_ppo_unpack_grads(norm2_accum) leaves
    native.check(tr.lib.dyros_ppo_adam_packed(C.byref(pk.desc), tr._p(tr.net.flat), tr._p(tr.net.grad), tr._p(tr.net.exp_avg),
                                              tr._p(tr.net.exp_avg_sq), 0.5, 0.5, 1, tr._p(tr.norm2), tr._p(tr.lr), tr._p(tr.opt_step),
                                              0.9, 0.999, 1e-8, 1e-5, 3e-6, 5000, tr._stream), "adam_packed")
    torch.cuda.synchronize()
    assert torch.allclose(tr.net.flat, ref["p"], rtol=1e-6, atol=1e-9) and torch.allclose(tr.net.exp_avg_sq, ref["v"], rtol=1e-5, atol=1e-12)
    assert float((tr.net.flat - ref["p"]).abs().max()) < 1e-8 and tr.lr[0].item() == pytest.approx(lr2[0].item(), rel=1e-6)
    w0 = tr.net.layers["critic_mlp.0"][0]
    assert torch.equal(pk.w0[1, :, :487].float(), w0.detach().to(torch.bfloat16).float()) and float(pk.w0[:, :, 487].abs().max()) == 0.0
    assert torch.equal(pk.bh[0, :13], tr.net.layers["mu"][1].detach()) and float(pk.wh[0, 13:].abs().max()) == 0.0
    env.close()


def test_checkpoint_round_trip_and_weight_dump(tmp_path):
    """PPOTrainer.save / restore (the fork's agent.save / restore, AG:97-103; format pinned in tests/test_ppo_formats.py) and
    the ./result weight dump of the play path."""
    env, tr = make_trainer(N=64, H=8, mb=256)
    tr.train_epoch()
    path = tr.save(str(tmp_path / "ckpt"))
    assert path.endswith("ckpt.pth")
    env2, tr2 = make_trainer(N=64, H=8, mb=256)
    assert float((tr2.net.flat - tr.net.flat).abs().max()) > 0
    tr2.restore(path)
    for a, b in ((tr.net.flat, tr2.net.flat), (tr.net.exp_avg, tr2.net.exp_avg), (tr.net.exp_avg_sq, tr2.net.exp_avg_sq),
                 (tr.net.logstd, tr2.net.logstd), (tr.packed.w0, tr2.packed.w0), (tr.packed.bh, tr2.packed.bh)):
        assert torch.equal(a, b)
    assert int(tr2.opt_step.item()) == int(tr.opt_step.item()) > 0 and tr2.epoch == 1
    assert tr2.lr[0].item() == pytest.approx(tr.lr[0].item(), rel=1e-6)
    files = tr.dump_weights_txt(str(tmp_path / "result"))
    assert len(files) == 13 and files[0].endswith("a2c_network_sigma.txt")
    w = np.loadtxt(str(tmp_path / "result" / "a2c_network_mu_weight.txt"))
    assert np.allclose(w, tr.net.layers["mu"][0].detach().cpu().numpy())
    env.close(); env2.close()


def test_on_device_ppo_learns_to_stay_up():
    """The whole stack end to end (fused env step with self-collision, packed bf16 networks, CUDA-graphed rollout and
    mini-epochs, the reference's hyper-parameters) at the benchmark's size: 60 epochs = 31 M env-steps, a few seconds.
    Episodes must get longer (profiles/r2_ppo_learning_curve.md: 108 -> ~380 policy steps by epoch 60)."""
    from isaacgymdyros_b200 import DyrosDynamicWalk, default_cfg
    from isaacgymdyros_b200.ppo import PPOConfig, PPOTrainer
    env = DyrosDynamicWalk(default_cfg(4096), DEV, use_cuda_graph=False)
    tr = PPOTrainer(env, PPOConfig())
    first = tr.train_epoch()
    for _ in range(58):
        tr.train_epoch()
    last = tr.train_epoch()
    assert all(math.isfinite(v) for v in last.values()), last
    assert first["episodes"] > 1000 and 60 < first["mean_length"] < 200, first
    assert last["mean_length"] > 2.0 * first["mean_length"] and last["mean_reward"] > 2.0 * first["mean_reward"], (first, last)
    assert bool(torch.isfinite(env.obs_buf).all())
    env.close()
