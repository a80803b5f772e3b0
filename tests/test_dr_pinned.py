"""Pins the domain-randomisation tables to the reference's own sampler: tests/golden/dr_samples.npz
was produced by the unmodified `gymutil.apply_random_samples` (tests/golden/make_dr_golden.py) with every uniform draw
recorded. CPU: the oracle's restatement against it; GPU: the reset kernel with the same draws injected."""
import os

import numpy as np
import pytest

from oracle import task_oracle as O
from tests.golden_util import GOLDEN_DIR

ARMATURE = np.array([0.614, 0.862, 1.09, 1.09, 1.09, 0.360, 0.614, 0.862, 1.09, 1.09, 1.09, 0.360, 0.078, 0.078, 0.078,
                     0.18, 0.18, 0.18, 0.18, 0.0032, 0.0032, 0.0032, 0.0032, 0.0032, 0.0032, 0.18, 0.18, 0.18, 0.18,
                     0.0032, 0.0032, 0.0032, 0.0032], np.float32)  # T:366-371
RTOL = 1e-6  # float32 formula vs numpy's float64 one rounded to the float32 property field


def golden():
    return np.load(os.path.join(GOLDEN_DIR, "dr_samples.npz"))


def test_oracle_dr_tables_match_reference_sampler():
    z = golden()
    for r in (0, 1):  # round 1 re-randomises the arrays round 0 already changed: no compounding
        d, a = O.dr_apply(z[f"r{r}/u_damping"], z[f"r{r}/u_armature"], base_armature=ARMATURE)
        assert np.allclose(d, z[f"r{r}/damping"], rtol=RTOL, atol=0)
        assert np.allclose(a, z[f"r{r}/armature"], rtol=RTOL, atol=0)
    assert z["r0/damping"].min() >= 0.1 and z["r0/damping"].max() <= 3.0          # 0.1 + U[0, 2.9]
    ratio = z["r1/armature"] / ARMATURE
    assert ratio.min() >= 0.8 - 1e-6 and ratio.max() <= 1.2 + 1e-6                 # og * U[0.8, 1.2], not compounded
    m = z["mass0"][None, :] * O.dr_mass_scale(z["u_mass"])
    assert np.allclose(m, z["mass"], rtol=RTOL, atol=0)


@pytest.mark.gpu
def test_cuda_reset_dr_tables_match_reference_sampler():
    """stage_reset_env with DyrosNoiseInjection.dr_u = the reference's recorded uniforms; two rounds on the same envs."""
    import torch
    from isaacgymdyros_b200.core import CoreConfig, DyrosCore
    z = golden()
    N = int(z["meta_N"])
    core = DyrosCore(N, "cuda:0", CoreConfig(randomize=True))
    dev = core.device
    ids = torch.arange(N, device=dev, dtype=torch.int64)
    zeros = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
    for r in (0, 1):
        u = torch.tensor(np.concatenate([z[f"r{r}/u_damping"], z[f"r{r}/u_armature"]], 1), device=dev).contiguous()
        core.set_noise_injection(dr_u=u, reset_f=zeros(N, 32), reset_i=zeros(N, 2, dt=torch.int64))
        core.task_t["randomize_buf"].fill_(3)  # VT:540-544: randomize_buf >= frequency
        core.reset_idx(ids)
        torch.cuda.synchronize()
        assert np.allclose(core.sim_t["dof_damping"].cpu().numpy(), z[f"r{r}/damping"], rtol=RTOL, atol=0)
        assert np.allclose(core.sim_t["dof_armature"].cpu().numpy(), z[f"r{r}/armature"], rtol=RTOL, atol=0)
        assert int(core.task_t["randomize_buf"].abs().sum()) == 0
    # VT:540-544: an env whose randomize_buf is below the frequency keeps its tables
    before = core.sim_t["dof_damping"].clone()
    core.set_noise_injection(dr_u=zeros(N, 66), reset_f=zeros(N, 32), reset_i=zeros(N, 2, dt=torch.int64))
    core.reset_idx(ids)
    torch.cuda.synchronize()
    assert torch.equal(core.sim_t["dof_damping"], before)
    core.close()
