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
