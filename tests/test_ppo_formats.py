"""On-disk formats of the reference's training stack written by isaacgymdyros_b200/ppo.py: the rl_games
checkpoint of the Dyros fork (a2c_common_dyros.py:550-585: model.state_dict(), optimizer_actor / optimizer_critic state
dicts) and the ./result/*.txt weight dump of the play path (torch_runner_dyros.py:140-150). rl_games itself is not
installed here, so the check is structural: a torch module tree with the names network_builder_dyros.py:78-103 /
models_dyros.py:17-20 give it and two torch.optim.Adam built as a2c_continuous_seperate.py:50-54 builds them must load
what we write with strict=True, and what they write must load back into the flat buffers."""
import numpy as np
import torch
from torch import nn

from isaacgymdyros_b200.ppo import FlatActorCritic, PPOConfig


class RefA2CNetwork(nn.Module):  # names and registration order of A2CBuilder.Network for `separate: True`, fixed sigma
    def __init__(self):
        super().__init__()
        mlp = lambda: nn.Sequential(nn.Linear(487, 256), nn.ELU(), nn.Linear(256, 256), nn.ELU())
        self.actor_cnn, self.critic_cnn = nn.Sequential(), nn.Sequential()
        self.actor_mlp, self.critic_mlp = mlp(), mlp()
        self.value = nn.Linear(256, 1)
        self.mu = nn.Linear(256, 13)
        self.sigma = nn.Parameter(torch.zeros(13), requires_grad=False)


class RefModel(nn.Module):  # ModelA2CContinuousLogStdDYROS.Network
    def __init__(self):
        super().__init__()
        self.a2c_network = RefA2CNetwork()


def ref_optimizers(model):
    a = list(model.a2c_network.actor_mlp.parameters()) + list(model.a2c_network.mu.parameters())
    c = list(model.a2c_network.critic_mlp.parameters()) + list(model.a2c_network.value.parameters())
    return torch.optim.Adam(a, lr=1e-5, eps=1e-8), torch.optim.Adam(c, lr=5e-4, eps=1e-8)


def test_model_and_optimizer_state_dicts_load_into_the_reference_layout_and_back():
    net = FlatActorCritic("cpu", PPOConfig())
    g = torch.Generator().manual_seed(0)
    net.flat.copy_(torch.randn(net.n, generator=g))
    net.exp_avg.copy_(torch.randn(net.n, generator=g)); net.exp_avg_sq.copy_(torch.rand(net.n, generator=g))
    net.logstd.fill_(-2.5)
    ref = RefModel()
    assert list(net.model_state_dict().keys()) == list(ref.state_dict().keys())          # names AND order (the txt dump follows it)
    ref.load_state_dict(net.model_state_dict(), strict=True)
    assert torch.equal(ref.a2c_network.actor_mlp[2].weight, net.layers["actor_mlp.1"][0])
    assert torch.equal(ref.a2c_network.value.bias, net.layers["value"][1])
    oa, oc = ref_optimizers(ref)
    oa.load_state_dict(net.optimizer_state_dict(True, 7, 9e-6))
    oc.load_state_dict(net.optimizer_state_dict(False, 7, 5e-4))
    p_mu = ref.a2c_network.mu.weight
    off = net.n_actor - 13 - 13 * 256
    assert torch.equal(oa.state[p_mu]["exp_avg"], net.exp_avg[off:off + 13 * 256].view(13, 256)) and float(oa.state[p_mu]["step"]) == 7
    assert oa.param_groups[0]["lr"] == 9e-6
    # ... one reference optimiser step each, then back into fresh flat buffers
    for p in list(ref.parameters()):
        if p.requires_grad:
            p.grad = torch.full_like(p, 0.01)
    oa.step(); oc.step()
    back = FlatActorCritic("cpu", PPOConfig())
    back.load_model_state_dict(ref.state_dict())
    assert back.load_optimizer_state_dict(True, oa.state_dict()) == 8 and back.load_optimizer_state_dict(False, oc.state_dict()) == 8
    assert torch.equal(back.layers["mu"][0], ref.a2c_network.mu.weight) and float(back.logstd[0]) == -2.5
    m, v = back._moment_views(back.layers["critic_mlp.0"][0])
    assert torch.equal(m, oc.state[ref.a2c_network.critic_mlp[0].weight]["exp_avg"])
    assert float((back.flat - net.flat).abs().max()) > 0                                  # the step moved the parameters


def test_weight_dump_names_follow_the_play_path(tmp_path):
    net = FlatActorCritic("cpu", PPOConfig())
    names = [k.replace(".", "_") + ".txt" for k in net.model_state_dict()]
    assert names[0] == "a2c_network_sigma.txt" and "a2c_network_actor_mlp_2_weight.txt" in names and names[-1] == "a2c_network_mu_bias.txt"
    for k, v in net.model_state_dict().items():                                           # np.savetxt round trip, as the controller reads it
        path = tmp_path / (k.replace(".", "_") + ".txt")
        np.savetxt(path, v.numpy())
        assert np.allclose(np.loadtxt(path).reshape(v.shape), v.numpy())
