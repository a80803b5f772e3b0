"""Self-collision (the reference creates the actor with collision filter 0, T:354, and `collision_true`
T:590 ends the episode on any non-foot contact): tables (model/selfcollision.py), oracle (oracle/selfcollision_oracle.py)
and the CUDA pass (k_self_collision through dyros_simulate / dyros_task_step)."""
import numpy as np
import pytest

from isaacgymdyros_b200.core import INIT_DOF_POS, self_collision_tables
from isaacgymdyros_b200.model import selfcollision as SC
from oracle.selfcollision_oracle import self_contact_forces
from tests.golden_util import load_assets


def tocabi():
    t, mocap, _ = load_assets()
    return t, mocap, self_collision_tables(t)


def folded_arm(t):
    q = np.array(INIT_DOF_POS)
    for n, v in zip(["L_Shoulder1_Joint", "L_Shoulder2_Joint", "L_Shoulder3_Joint", "L_Armlink_Joint", "L_Elbow_Joint", "L_Forearm_Joint"],
                    [0.66, 0.08, 1.65, -2.73, 1.86, -2.72]):
        q[t.dof_names.index(n)] = v
    return q


def test_tables_known_answers():
    t, _, sc = tocabi()
    assert sc.num_shapes == 61 and (sc.shape_kind == 0).sum() == 25 and (sc.shape_kind == 1).sum() == 36   # XML:99-353
    assert len(sc.sample) == 25 * 8 + 36 * 3
    parent = np.asarray(t.link_parent)
    for i, j in sc.pairs:
        assert i < j and parent[j] != i and parent[i] != j                       # links joined by a joint never collide
    # the sole box recovered from its 8 corners: 0.30 x 0.17 x 0.007 (XML:148)
    foot = [k for k in range(sc.num_shapes) if t.body_names[sc.shape_body[k]] == "L_Foot_Link" and sc.shape_kind[k] == 0][0]
    assert sorted(np.round(sc.shape_size[foot], 4)) == [0.0035, 0.085, 0.15]
    assert np.allclose(sc.shape_rot[foot].reshape(3, 3) @ sc.shape_rot[foot].reshape(3, 3).T, np.eye(3), atol=1e-9)


def test_sdf_known_answers():
    d, g = SC.sdf(SC.KIND_BOX, np.array([1.0, 2.0, 3.0]), np.array([[2.0, 0, 0], [0.5, 0, 0], [2.0, 3.0, 0], [0, 0, -2.9]]))
    assert np.allclose(d, [1.0, -0.5, np.sqrt(2), -0.1]) and np.allclose(g[0], [1, 0, 0]) and np.allclose(g[3], [0, 0, -1])
    d, g = SC.sdf(SC.KIND_CYL, np.array([1.0, 2.0, 0.0]), np.array([[3.0, 0, 0], [0, 0, 2.5], [0.5, 0, 0], [2.0, 0, 3.0]]))
    assert np.allclose(d, [2.0, 0.5, -0.5, np.sqrt(2)]) and np.allclose(g[1], [0, 0, 1]) and np.allclose(g[2], [1, 0, 0])


def test_gait_is_collision_free_and_contortions_are_not():
    t, mocap, sc = tocabi()
    for row in range(0, 3600, 100):                                              # the reference's walking cycle (T:112-116)
        Rw, pw = SC.link_fk(t, mocap[row, 1:34])
        assert np.abs(self_contact_forces(sc, Rw, pw, 2e5, 2e4)).max() == 0.0
    Rw, pw = SC.link_fk(t, np.array(INIT_DOF_POS))
    assert np.abs(self_contact_forces(sc, Rw, pw, 2e5, 2e4)).max() == 0.0
    # left arm folded across the chest: the forearm / wrist end up inside the torso boxes
    q = folded_arm(t)
    Rw, pw = SC.link_fk(t, q)
    F = self_contact_forces(sc, Rw, pw, 2e5, 2e4)
    hit = [t.body_names[b] for b in np.nonzero(np.linalg.norm(F, axis=1) > 1.0)[0]]
    assert "Upperbody_Link" in hit and "L_Forearm_Link" in hit, hit
    assert np.abs(F.sum(0)).max() < 1e-6                                         # action = reaction


@pytest.mark.gpu
@pytest.mark.parametrize("program,caps", [("roles", None), ("lanes", None), ("roles", "5")])
def test_cuda_self_collision_matches_oracle(program, caps, monkeypatch):
    """Random contorted poses through dyros_simulate: self_contact_force (and its share of net_contact_force) against the
    oracle evaluated on the poses the sub-step starts from; both physics programs export the same link poses. `caps`
    shrinks the kernel's hit list so that these poses overflow it (the fallback must give the same forces)."""
    import torch
    if caps:
        monkeypatch.setenv("DYROS_SC_TEST_CAPS", caps)
    from isaacgymdyros_b200.core import CoreConfig, DyrosCore
    from oracle.physics_oracle import PhysicsOracle
    from tests.physics_util import oracle_params
    t, _, sc = tocabi()
    N = 96 if caps is None else 37                                               # (37: a partly filled CTA of the pass)
    rng = np.random.default_rng(11)
    cfg = CoreConfig(physics_program=program)
    core = DyrosCore(N, "cuda:0", cfg)
    q = np.clip(np.array(INIT_DOF_POS) + rng.normal(0, 0.45, (N, 33)), t.dof_lower, t.dof_upper).astype(np.float32)
    q[: N // 4] = np.array(INIT_DOF_POS, np.float32)                            # a quarter stands in the initial pose
    root = np.zeros((N, 13), np.float32)
    root[:, 2] = 3.0                                                             # in the air: no ground contact
    quat = rng.normal(0, 1, (N, 4)); root[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    core.sim_t["root_states"].copy_(torch.tensor(root))
    ds = core.sim_t["dof_state"].view(N, 33, 2)
    ds[:, :, 0] = torch.tensor(q); ds[:, :, 1] = 0
    core.simulate()
    torch.cuda.synchronize()
    got = core.sim_t["self_contact_force"].view(N, 38, 3).cpu().numpy().astype(np.float64)
    net = core.sim_t["net_contact_force"].view(N, 38, 3).cpu().numpy().astype(np.float64)
    o = PhysicsOracle(t, oracle_params(cfg))
    _, Rw, pw = o.kinematics(root.astype(np.float64), q.astype(np.float64))
    want = np.stack([self_contact_forces(sc, [R[n] for R in Rw], [p[n] for p in pw], cfg.penalty_stiffness, cfg.penalty_max_force)
                     for n in range(N)])
    hit = np.linalg.norm(want, axis=2).max(1) > 1.0
    assert 0.3 < hit.mean() < 1.0 and not hit[: N // 4].any()
    # penalty forces are stiffness x depth: 2e5 N/m turns float32 pose rounding (1e-7 m) into ~0.05 N
    assert np.abs(got - want).max() < 2e-3 * np.abs(want).max() + 0.5
    assert np.array_equal(net, got)                                              # airborne: nothing else touches
    assert (np.linalg.norm(got, axis=2).max(1) > 1.0).tolist() == hit.tolist()
    core.close()


def test_batched_oracle_equals_the_loop_oracle():
    """oracle/selfcollision_oracle.py states the computation twice: per env with explicit loops (the restatement the
    kernel is compared with) and vectorised over envs (what EnvOracle and bench.py's CPU arms run). Same forces."""
    from oracle.selfcollision_oracle import self_contact_forces_batch
    from tests.test_humanoid_generality import humanoid
    rng = np.random.default_rng(5)
    for t, q0, sig in ((tocabi()[0], np.array(INIT_DOF_POS), 0.5), (humanoid(), None, None)):
        sc = self_collision_tables(t)
        N = 12
        qs = [np.clip(q0 + rng.normal(0, sig, len(q0)), t.dof_lower, t.dof_upper) if q0 is not None
              else rng.uniform(t.dof_lower, t.dof_upper) for _ in range(N)]
        fks = [SC.link_fk(t, q) for q in qs]
        want = np.stack([self_contact_forces(sc, Rw, pw, 2e5, 2e4) for Rw, pw in fks])
        Rw = [np.stack([fk[0][l] for fk in fks]) for l in range(t.num_links)]
        pw = [np.stack([fk[1][l] for fk in fks]) for l in range(t.num_links)]
        got = self_contact_forces_batch(sc, Rw, pw, 2e5, 2e4)
        assert np.abs(want).max() > 100.0 and np.allclose(got[:, : want.shape[1]], want, rtol=1e-9, atol=1e-6)


def test_humanoid_capsule_tables_known_answers():
    """The stock Humanoid (assets/mjcf/nv_humanoid.xml; tasks/humanoid.py creates its actor with filter 0 as well): capsules
    and spheres as shape kind 2."""
    from tests.test_humanoid_generality import humanoid
    t = humanoid()
    sc = self_collision_tables(t)
    assert sc.num_shapes == 19 and (sc.shape_kind == SC.KIND_CAP).all() and np.diff(sc.shape_sample0).max() <= 8
    head = [k for k in range(19) if t.body_names[sc.shape_body[k]] == "head"][0]          # <geom name="head" type="sphere" size=".09">
    assert sc.shape_size[head][0] == pytest.approx(0.09) and sc.shape_size[head][1] == 0.0 and sc.shape_sample0[head + 1] - sc.shape_sample0[head] == 1
    d, g = SC.sdf(SC.KIND_CAP, np.array([0.05, 0.2, 0.0]), np.array([[0.1, 0, 0.1], [0, 0, 0.3], [0.03, 0, -0.5]]))
    assert np.allclose(d, [0.05, 0.05, np.hypot(0.03, 0.3) - 0.05]) and np.allclose(g[0], [1, 0, 0]) and np.allclose(g[1], [0, 0, 1])
    Rw, pw = SC.link_fk(t, np.zeros(t.num_dofs))
    assert np.abs(self_contact_forces(sc, Rw, pw, 2e5, 2e4)).max() == 0.0                   # the zero pose is collision-free
    rng = np.random.default_rng(1)
    hit = 0
    for _ in range(20):
        Rw, pw = SC.link_fk(t, rng.uniform(t.dof_lower, t.dof_upper))
        F = self_contact_forces(sc, Rw, pw, 2e5, 2e4)
        assert np.abs(F.sum(0)).max() < 1e-6
        hit += np.linalg.norm(F, axis=1).max() > 1.0
    assert 5 <= hit <= 19                                                                  # poses anywhere in the limits often self-intersect


@pytest.mark.gpu
def test_cuda_humanoid_self_collision_matches_oracle():
    import torch
    from isaacgymdyros_b200.core import CoreConfig, DyrosCore
    from oracle.physics_oracle import PhysicsOracle
    from tests.physics_util import oracle_params
    from tests.test_humanoid_generality import HUMANOID_CFG, humanoid
    t = humanoid()
    sc = self_collision_tables(t)
    N = 48
    rng = np.random.default_rng(3)
    cfg = CoreConfig(**HUMANOID_CFG)
    core = DyrosCore(N, "cuda:0", cfg, tables=t, with_task=False)
    q = rng.uniform(t.dof_lower, t.dof_upper, (N, t.num_dofs)).astype(np.float32)
    q[: N // 4] = 0.0
    root = np.zeros((N, 13), np.float32)
    root[:, 2] = 4.0
    quat = rng.normal(0, 1, (N, 4)); root[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    core.sim_t["root_states"].copy_(torch.tensor(root))
    ds = core.sim_t["dof_state"].view(N, t.num_dofs, 2)
    ds[:, :, 0] = torch.tensor(q); ds[:, :, 1] = 0
    core.simulate()
    torch.cuda.synchronize()
    got = core.sim_t["self_contact_force"].view(N, t.num_bodies, 3).cpu().numpy().astype(np.float64)
    o = PhysicsOracle(t, oracle_params(cfg), solver_bodies=cfg.solver_bodies)
    _, Rw, pw = o.kinematics(root.astype(np.float64), q.astype(np.float64))
    want = np.stack([self_contact_forces(sc, [R[n] for R in Rw], [p[n] for p in pw], cfg.penalty_stiffness, cfg.penalty_max_force)
                     for n in range(N)])
    hit = np.linalg.norm(want, axis=2).max(1) > 1.0
    assert 0.2 < hit.mean() < 1.0 and not hit[: N // 4].any()
    assert np.abs(got - want).max() < 2e-3 * np.abs(want).max() + 0.5
    core.close()


@pytest.mark.gpu
def test_self_collision_terminates_the_episode_in_the_fused_step():
    """T:590: a non-foot body in (self-)contact resets the env in the same step; without the tables it does not."""
    import torch
    from isaacgymdyros_b200.core import CoreConfig, DyrosCore
    t, _, _ = tocabi()
    res = {}
    for on in (True, False):
        core = DyrosCore(8, "cuda:0", CoreConfig(self_collision=on, perturb=False))
        ds = core.sim_t["dof_state"].view(8, 33, 2)
        ds[0, :, 0] = torch.tensor(folded_arm(t), dtype=torch.float32)
        core.task_t["reset_buf"].zero_()
        core.step(torch.zeros(8, 13, device="cuda:0"))
        torch.cuda.synchronize()
        res[on] = core.task_t["reset_buf"].cpu().tolist()
        assert core.step_launches() == (3 if on else 2)
        core.close()
    assert res[True][0] == 1 and sum(res[True][1:]) == 0
    assert res[False][0] == 0
