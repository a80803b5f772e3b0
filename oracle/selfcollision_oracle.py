"""TEST INFRASTRUCTURE ONLY -- numpy (float64) restatement of the self-collision detection the CUDA kernel
k_self_collision performs (csrc/task_kernels.cu), on the tables of isaacgymdyros_b200/model/selfcollision.py.
Reference behaviour it stands for: PhysX self-collision of an actor created with collision filter 0
(tasks/dyros_dynamic_walk.py:354), feeding the net contact forces that `collision_true` tests (T:590, T:937).
PhysX is closed source and absent: parity with it is unpinned (as for the rest of gym.simulate); what the tests pin is
kernel == this restatement, plus geometric known answers.

Written shape by shape with explicit loops (no shared code with the product's vectorised table builder beyond the
tables themselves)."""
import numpy as np

KIND_BOX, KIND_CYL, KIND_CAP = 0, 1, 2


def sdf_point(kind, size, x):
    """Signed distance and outward unit gradient of ONE point x (3,) in the shape's frame."""
    if kind == KIND_BOX:
        q = np.abs(x) - size
        out = np.maximum(q, 0.0)
        n = np.linalg.norm(out)
        if n > 0:
            return n, out / n * np.where(x < 0, -1.0, 1.0)
        k = int(np.argmax(q))
        g = np.zeros(3)
        g[k] = -1.0 if x[k] < 0 else 1.0
        return float(q[k]), g
    r, h = size[0], size[1]
    if kind == KIND_CAP:                                       # capsule (sphere when h = 0): segment [-h, h] on z
        q = np.array([x[0], x[1], x[2] - min(max(x[2], -h), h)])
        n = np.linalg.norm(q)
        return n - r, (q / n if n > 0 else np.array([1.0, 0.0, 0.0]))
    rho = np.hypot(x[0], x[1])
    er = np.array([x[0], x[1], 0.0]) / max(rho, 1e-30)
    ez = np.array([0.0, 0.0, -1.0 if x[2] < 0 else 1.0])
    qr, qz = rho - r, abs(x[2]) - h
    o = np.array([max(qr, 0.0), max(qz, 0.0)])
    n = np.linalg.norm(o)
    if n > 0:
        return n, (o[0] * er + o[1] * ez) / n
    return (qr, er) if qr > qz else (qz, ez)


def self_contact_forces(sc, Rw, pw, stiffness, fmax):
    """Net self-contact force per body, (nb, 3) world axes, for ONE env: Rw[l] (3,3), pw[l] (3,) world pose of link l."""
    nb = int(sc.shape_body.max()) + 1 if len(sc.shape_body) else 0
    F = np.zeros((max(nb, 1), 3))
    for i, j in sc.pairs:
        ci = Rw[i] @ sc.link_sphere[i, :3] + pw[i]
        cj = Rw[j] @ sc.link_sphere[j, :3] + pw[j]
        if np.linalg.norm(ci - cj) >= sc.link_sphere[i, 3] + sc.link_sphere[j, 3]:
            continue
        for (la, lb) in ((i, j), (j, i)):                      # samples of the shapes of `la` against the shapes of `lb`
            for sa in range(sc.link_shape0[la], sc.link_shape0[la + 1]):
                for k in range(sc.shape_sample0[sa], sc.shape_sample0[sa + 1]):
                    c = Rw[la] @ sc.sample[k, :3] + pw[la]
                    rho = sc.sample[k, 3]
                    for sb in range(sc.link_shape0[lb], sc.link_shape0[lb + 1]):
                        Rs = sc.shape_rot[sb].reshape(3, 3)
                        x = Rs.T @ (Rw[lb].T @ (c - pw[lb]) - sc.shape_center[sb])
                        d, g = sdf_point(int(sc.shape_kind[sb]), sc.shape_size[sb], x)
                        depth = rho - d
                        if depth > 0:
                            f = min(stiffness * depth, fmax) * (Rw[lb] @ (Rs @ g))   # pushes the sample's body out of B
                            F[sc.shape_body[sa]] += f
                            F[sc.shape_body[sb]] -= f
    return F


# ---------------------------------------------------------------------------------------------------------------------
# The same computation for many envs at once (numpy over the envs that pass a pair's broad phase): what EnvOracle and
# bench.py's CPU arms use, so that the self-collision pass does not dominate the CPU baseline. Checked against the loop
# above in tests/test_self_collision.py.
def _sdf_many(kind, size, x):
    """Signed distance and outward unit gradient of points x (M, 3) in the shape's frame."""
    if kind == KIND_BOX:
        q = np.abs(x) - size
        out = np.maximum(q, 0.0)
        n = np.linalg.norm(out, axis=1)
        sgn = np.where(x < 0, -1.0, 1.0)
        g_out = out / np.maximum(n, 1e-300)[:, None] * sgn
        k = np.argmax(q, axis=1)
        g_in = np.zeros_like(x)
        g_in[np.arange(len(x)), k] = sgn[np.arange(len(x)), k]
        d_in = q[np.arange(len(x)), k]
        outside = n > 0
        return np.where(outside, n, d_in), np.where(outside[:, None], g_out, g_in)
    r, h = size[0], size[1]
    if kind == KIND_CAP:
        q = x.copy()
        q[:, 2] -= np.clip(x[:, 2], -h, h)
        n = np.linalg.norm(q, axis=1)
        g = np.where((n > 0)[:, None], q / np.maximum(n, 1e-300)[:, None], np.array([1.0, 0.0, 0.0]))
        return n - r, g
    rho = np.hypot(x[:, 0], x[:, 1])
    er = np.stack([x[:, 0], x[:, 1], np.zeros(len(x))], 1) / np.maximum(rho, 1e-30)[:, None]
    ez = np.stack([np.zeros(len(x)), np.zeros(len(x)), np.where(x[:, 2] < 0, -1.0, 1.0)], 1)
    qr, qz = rho - r, np.abs(x[:, 2]) - h
    o0, o1 = np.maximum(qr, 0.0), np.maximum(qz, 0.0)
    n = np.hypot(o0, o1)
    g_out = (o0[:, None] * er + o1[:, None] * ez) / np.maximum(n, 1e-300)[:, None]
    inside_r = qr > qz
    outside = n > 0
    d = np.where(outside, n, np.where(inside_r, qr, qz))
    g = np.where(outside[:, None], g_out, np.where(inside_r[:, None], er, ez))
    return d, g


def self_contact_forces_batch(sc, Rw, pw, stiffness, fmax):
    """Rw[l] (N,3,3), pw[l] (N,3): world poses of every link for N envs -> (N, nb, 3)."""
    N = pw[0].shape[0]
    nb = int(sc.shape_body.max()) + 1 if len(sc.shape_body) else 1
    F = np.zeros((N, nb, 3))
    if not len(sc.pairs):
        return F
    nl = len(Rw)
    C = np.stack([np.einsum("nij,j->ni", Rw[l], sc.link_sphere[l, :3]) + pw[l] for l in range(nl)], 0)   # (nl, N, 3)
    P = np.asarray(sc.pairs)
    overlap = np.linalg.norm(C[P[:, 0]] - C[P[:, 1]], axis=2) < (sc.link_sphere[P[:, 0], 3] + sc.link_sphere[P[:, 1], 3])[:, None]
    for pi in np.nonzero(overlap.any(axis=1))[0]:                                            # pairs some env has near
        i, j = int(P[pi, 0]), int(P[pi, 1])
        near = np.nonzero(overlap[pi])[0]
        for la, lb in ((i, j), (j, i)):
            Ra, pa, Rb, pb = Rw[la][near], pw[la][near], Rw[lb][near], pw[lb][near]
            k0, k1 = sc.shape_sample0[sc.link_shape0[la]], sc.shape_sample0[sc.link_shape0[la + 1]]
            if k1 == k0:
                continue
            smp = sc.sample[k0:k1]                                                            # (S, 4) all samples of link la
            body_a = np.repeat(sc.shape_body[sc.link_shape0[la]:sc.link_shape0[la + 1]],
                               np.diff(sc.shape_sample0[sc.link_shape0[la]:sc.link_shape0[la + 1] + 1]))
            c = np.einsum("nij,sj->nsi", Ra, smp[:, :3]) + pa[:, None, :]                    # (E, S, 3) world
            cl = np.einsum("nji,nsj->nsi", Rb, c - pb[:, None, :])                           # in link lb's frame
            E, S = cl.shape[0], cl.shape[1]
            for sb in range(sc.link_shape0[lb], sc.link_shape0[lb + 1]):
                Rs = sc.shape_rot[sb].reshape(3, 3)
                d, g = _sdf_many(int(sc.shape_kind[sb]), sc.shape_size[sb], ((cl - sc.shape_center[sb]) @ Rs).reshape(E * S, 3))
                depth = (smp[None, :, 3] - d.reshape(E, S))
                en, sn = np.nonzero(depth > 0)
                if not len(en):
                    continue
                gw = np.einsum("nij,nj->ni", Rb[en], g.reshape(E, S, 3)[en, sn] @ Rs.T)
                f = np.minimum(stiffness * depth[en, sn], fmax)[:, None] * gw
                np.add.at(F, (near[en], body_a[sn]), f)
                np.add.at(F, (near[en], np.full(len(en), sc.shape_body[sb])), -f)
    return F
