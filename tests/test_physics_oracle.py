"""Physical invariants that pin oracle/physics_oracle.py (the fp64 statement of OUR simulate model; PhysX
itself is absent from the reference checkout, so parity with it is unpinned -- see the oracle's header)."""
import numpy as np
import pytest

from oracle.physics_oracle import PhysicsOracle, PhysParams
from tests.golden_util import load_assets
from isaacgymdyros_b200.core import ARMATURE, INIT_DOF_POS, KP, KV

G = 9.81


@pytest.fixture(scope="module")
def tables():
    return load_assets()[0]


def base_state(N, z, rng=None):
    root = np.zeros((N, 13))
    root[:, 2] = z
    root[:, 6] = 1.0
    q = np.tile(np.array(INIT_DOF_POS), (N, 1))
    qd = np.zeros((N, 33))
    if rng is not None:
        quat = rng.normal(0, 1, (N, 4))
        root[:, 3:7] = quat / np.linalg.norm(quat, axis=-1, keepdims=True)
        root[:, 7:13] = rng.normal(0, 0.5, (N, 6))
        q = q + rng.normal(0, 0.2, (N, 33))
        qd = rng.normal(0, 0.5, (N, 33))
    return root, q, qd


def test_mass_matrix_spd_and_total_mass(tables):
    o = PhysicsOracle(tables)
    rng = np.random.default_rng(0)
    root, q, qd = base_state(3, 2.0, rng)
    ones = np.ones((3, 38))
    *_, d = o.substep(root, q, qd, np.zeros((3, 33)), np.zeros((3, 33)), np.zeros((3, 33)), ones)
    M = d["M"]
    assert np.allclose(M, np.swapaxes(M, -1, -2), atol=1e-10)
    assert (np.linalg.eigvalsh(M) > 0).all()
    assert np.allclose(M[:, 3, 3], 104.48712) and np.allclose(M[:, 4, 4], 104.48712)  # translational block = total mass


def test_free_fall_momentum_and_energy(tables):
    o = PhysicsOracle(tables)
    rng = np.random.default_rng(1)
    N, steps = 4, 25
    root, q, qd = base_state(N, 3.0, rng)
    ms = rng.uniform(0.8, 1.2, (N, 38))
    arm = np.zeros((N, 33))  # rotor inertia carries kinetic energy the link-level sum does not see: off here
    mass = (tables.body_inertia[:, 0][None, :] * ms).sum(1)
    l0, a0, k0, p0 = o.momentum_energy(root, q, qd, ms)
    com0 = None
    for _ in range(steps):
        root, q, qd, cf, _d = o.substep(root, q, qd, np.zeros((N, 33)), np.zeros((N, 33)), arm + 1e-3, ms)
        assert not cf.any()
    l1, a1, k1, p1 = o.momentum_energy(root, q, qd, ms)
    T = steps * o.p.dt
    want = np.zeros((N, 3))
    want[:, 2] = -mass * G * T
    assert np.allclose(l1 - l0, want, rtol=0, atol=2e-2)  # first-order integrator: O(dt) per unit time
    assert np.allclose((k1 + p1), (k0 + p0), rtol=2e-3)


def test_zero_gravity_internal_torques_conserve_momentum(tables):
    """Joint torques, damping and rotor inertia are internal: total momentum is conserved up to the
    first-order discretisation error, which must halve when dt halves."""
    errs = []
    for dt, steps in [(0.002, 20), (0.001, 40)]:
        o = PhysicsOracle(tables, PhysParams(gravity=(0, 0, 0), vel_limit=1e9, dt=dt))
        rng = np.random.default_rng(2)
        N = 3
        root, q, qd = base_state(N, 3.0, rng)
        q = np.clip(q, tables.dof_lower + 0.3, tables.dof_upper - 0.3)
        ones = np.ones((N, 38))
        arm = np.tile(np.array(ARMATURE), (N, 1))
        tau = rng.normal(0, 1, (N, 33)) * np.array(ARMATURE) * 30
        l0, a0, *_ = o.momentum_energy(root, q, qd, ones)
        for _ in range(steps):
            root, q, qd, _cf, _d = o.substep(root, q, qd, tau, np.full((N, 33), 0.1), arm, ones)
        l1, a1, *_ = o.momentum_energy(root, q, qd, ones)
        errs.append((np.abs(l1 - l0).max(), np.abs(a1 - a0).max()))
    assert errs[0][0] < 0.1 and errs[0][1] < 0.3            # ~1e-3 of the momenta involved
    assert 1.8 < errs[0][0] / errs[1][0] < 2.2 and 1.8 < errs[0][1] / errs[1][1] < 2.2


def test_push_changes_linear_momentum_by_impulse(tables):
    o = PhysicsOracle(tables, PhysParams(gravity=(0, 0, 0)))
    N = 2
    root, q, qd = base_state(N, 3.0)
    ones = np.ones((N, 38))
    push = np.array([[100.0, -50.0, 0.0], [0.0, 30.0, 10.0]])
    l0, *_ = o.momentum_energy(root, q, qd, ones)
    root, q, qd, _cf, _d = o.substep(root, q, qd, np.zeros((N, 33)), np.zeros((N, 33)), np.ones((N, 33)), ones, push=push)
    l1, *_ = o.momentum_energy(root, q, qd, ones)
    assert np.allclose(l1 - l0, push * o.p.dt, atol=1e-6)


def test_standing_contact_carries_the_weight_on_the_feet_only(tables):
    """Reset pose (T:95-100, height 0.93): the soles start 1.47 mm above the plane, i.e. inside
    contact_offset; with the joints held by a stiff PD the ground reaction settles at m*g on bodies 8 and 16."""
    o = PhysicsOracle(tables)
    N = 1
    root, q, qd = base_state(N, 0.93)
    X, Rw, pw = o.kinematics(root, q)
    assert abs((pw[6][0, 2] - 0.1585) - 0.00147) < 2e-5
    q0 = q.copy()
    kp, kv = np.array(KP), np.array(KV)
    ones = np.ones((N, 38))
    arm = np.tile(np.array(ARMATURE), (N, 1))
    fz = []
    for s in range(400):
        tau = kp * (q0 - q) - kv * qd
        root, q, qd, cf, _d = o.substep(root, q, qd, tau, np.full((N, 33), 0.1), arm, ones)
        if s >= 300:
            fz.append(cf[0, :, 2].sum())
            touching = {b for b in range(38) if np.abs(cf[0, b]).sum() > 0}
            assert touching == {8, 16}
    assert abs(np.mean(fz) - 104.48712 * G) < 0.02 * 104.48712 * G
    assert abs(root[0, 2] - 0.93) < 0.01 and abs(root[0, 3:6]).max() < 0.02


def test_joint_velocity_cap_and_limits(tables):
    o = PhysicsOracle(tables)
    N = 1
    root, q, qd = base_state(N, 3.0)
    tau = np.zeros((N, 33))
    tau[:, 3] = 5000.0
    ones = np.ones((N, 38))
    for _ in range(30):
        root, q, qd, _cf, _d = o.substep(root, q, qd, tau, np.full((N, 33), 0.1), np.tile(np.array(ARMATURE), (N, 1)), ones)
    assert np.abs(qd).max() <= 4.03 + 1e-12 and abs(qd[0, 3] - 4.03) < 1e-9
    assert (q <= tables.dof_upper + 1e-12).all() and (q >= tables.dof_lower - 1e-12).all()
