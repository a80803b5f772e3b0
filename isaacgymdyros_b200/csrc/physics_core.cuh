// One sub-step of gym.simulate for ONE env, written as the program of one of DYROS_LANES cooperating lanes.
//
// Replaces the body of `gym.simulate` (reference call site tasks/dyros_dynamic_walk.py:525; the reference's
// implementation is closed-source PhysX) for a floating-base tree of revolute joints: articulated-body
// forward dynamics (Featherstone ABA, three passes over a branch-parallel link schedule) with implicit joint
// damping and rotor inertia, penalty ground contact for all collision primitives except the sole corners,
// and a fixed-sweep projected Gauss-Seidel solve of the sole-corner contacts in the 6-D space of each foot
// link using the exact articulated inverse inertia (O(n) recursion down the leg chains). The model and every
// formula are stated independently (dense, fp64) in oracle/physics_oracle.py; see DESIGN.md section 4.
//
// The lanes of an env talk through the env's scratch block `sm` (shared memory on the GPU) and meet at
// sync() points; the same source is compiled for the host by tests/native/hostemu.cpp (4 threads + a barrier).
#pragma once
#include "internal.h"
#include "phys_math.cuh"

namespace dyros {

// per-link scratch layout (floats)
constexpr int LS_E = 0;    // 9  parent->link rotation (base: base->world rotation)
constexpr int LS_V = 9;    // 6  link velocity, later the impulse response dv
constexpr int LS_A = 15;   // 28 pass1: inertia(10) pA(6) world pose(12) | pass2: contribution to parent IA(21) pA(6) | pass3: a'(6)
constexpr int LS_U = 43;   // 6  U = IA S (base: predicted velocity v0*)
constexpr int LS_SC = 49;  // 4  [qd -> qd*, tau -> u, damping -> 1/D, armature -> S^T dp]
constexpr int LS = 53;
constexpr int A_INERTIA = 0, A_PA = 10, A_POSE = 16, A_CIA = 0, A_CPA = 21, A_ACC = 0, A_OM0 = 6;
// per-env extra scratch
constexpr int X_FOOTPOSE = 0;                     // MAX_FEET * 12
constexpr int X_Z = X_FOOTPOSE + MAX_FEET * 12;   // MAX_FEET * 6  base velocity change caused by a foot's sweep
constexpr int X_PD = X_Z + MAX_FEET * 6;          // MAX_FEET * 6  impulse arriving at the base from a foot
constexpr int X_ROWS = X_PD + MAX_FEET * 6;       // MAX_FEET * MAX_ACTIVE_PTS * 3 * 7  (Om J^T (6), 1 / (J Om J^T))
constexpr int X_SIZE = X_ROWS + MAX_FEET * MAX_ACTIVE_PTS * 3 * 7;

HD int env_scratch_floats(int nl) { return nl * LS + X_SIZE; }

// The hot model tables are read from the staged copy `hot` (shared memory on the GPU) through word offsets.
#define HI(field, idx) (reinterpret_cast<const int*>(hot)[m.o_##field + (idx)])
#define HF(field, idx) (hot[m.o_##field + (idx)])
#define HF3(field, idx) ld3_f(hot + m.o_##field + 3 * (idx))

// global-memory views of one env (all device pointers on the GPU, host pointers in the emulation)
struct EnvIO {
  float* root;             // 13
  float* dof_state;        // nd*2
  const float* tau;        // nd
  const float* damping;    // nd
  const float* armature;   // nd
  const float* mass_scale; // nb
  float* contact;          // nb*3
  const float* push;       // 3 or NULL
  const float* rb_force;   // nb*3 or NULL
  const float* rb_torque;  // nb*3 or NULL
  bool live;               // false: padding lane group, no global writes
};

// ---- penalty ground contact of one location on a link (oracle: PhysicsOracle._external_wrench.add_point)
HD void penalty_point(const SimParams& p, const M3& Rw, SV v, V3 xs, real depth, float* cf, bool live, SV& fext) {
  if (!(depth > 0)) return;
  V3 vel_w = mul(Rw, v.v + cross(v.w, xs));
  real fn = p.pen_k * depth - p.pen_c * vel_w.z;
  fn = fn < 0 ? 0 : (fn > p.pen_fmax ? p.pen_fmax : fn);
  real speed = sqrt(vel_w.x * vel_w.x + vel_w.y * vel_w.y);
  real lim = p.mu * fn / (speed > (real)1e-6 ? speed : (real)1e-6);
  real coef = p.pen_c < lim ? p.pen_c : lim;
  V3 Fw = v3(-coef * vel_w.x, -coef * vel_w.y, fn);
  if (live) {
    cf[0] += (float)Fw.x;
    cf[1] += (float)Fw.y;
    cf[2] += (float)Fw.z;
  }
  V3 fl = mulT(Rw, Fw);
  fext.w = fext.w + cross(xs, fl);
  fext.v = fext.v + fl;
}

// Link inertia parameters of link i from the per-body mass scales (P0). Writes A_INERTIA.
HD void link_inertia(const EnvIO& io, real* L, const float* hot, const DevModel& m, int i) {
  real par[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) par[k] = 0;
  for (int bi = HI(body_start, i); bi < HI(body_start, i + 1); ++bi) {
    int b = HI(bodies, bi);
    real sc = io.mass_scale[b];
#pragma unroll
    for (int k = 0; k < 10; ++k) par[k] += sc * HF(body_inertia, b * 10 + k);
  }
#pragma unroll
  for (int k = 0; k < 10; ++k) L[LS_A + A_INERTIA + k] = par[k];
}

// Bias force and external wrench of link i (pass 1). Reads A_INERTIA, writes A_PA.
HD void link_forces(const EnvIO& io, real* L, const float* hot, const DevModel& m, const SimParams& p, int i, const M3& Rw,
                    V3 pw, SV v) {
  real* A = L + LS_A;
  SV fext = sv_zero();
  V3 nrm = v3(Rw.a[6], Rw.a[7], Rw.a[8]);  // world z in link coordinates
  for (int bi = HI(body_start, i); bi < HI(body_start, i + 1); ++bi) {
    int b = HI(bodies, bi);
    if (io.live) {
      io.contact[3 * b] = 0.f;
      io.contact[3 * b + 1] = 0.f;
      io.contact[3 * b + 2] = 0.f;
    }
    bool has_push = io.push && b == 0;
    if (has_push || io.rb_force) {  // world wrench at the body's centre of mass (tensors.rst.txt:322-335)
      V3 F = v3(0, 0, 0), T = v3(0, 0, 0);
      if (has_push) F = ld3_f(io.push);
      if (io.rb_force) {
        F = F + ld3_f(io.rb_force + 3 * b);
        T = ld3_f(io.rb_torque + 3 * b);
      }
      real sc = io.mass_scale[b];
      real mb = sc * HF(body_inertia, b * 10);
      real inv = 1 / (mb > (real)1e-30 ? mb : (real)1e-30);
      V3 com = v3(sc * HF(body_inertia, b * 10 + 1) * inv, sc * HF(body_inertia, b * 10 + 2) * inv,
                  sc * HF(body_inertia, b * 10 + 3) * inv);
      V3 fl = mulT(Rw, F);
      fext.w = fext.w + cross(com, fl) + mulT(Rw, T);
      fext.v = fext.v + fl;
    }
  }
  const bool near_ground = pw.z < HF(reach, i);  // nothing of this link can reach z = 0 otherwise
  for (int k = near_ground ? HI(pt_start, i) : 0; k < (near_ground ? HI(pt_start, i + 1) : 0); ++k) {
    V3 x = ld3_f(m.pt_pos + 3 * k);
    real rad = m.pt_radius[k];
    real z = pw.z + dot(nrm, x);
    penalty_point(p, Rw, v, x - rad * nrm, rad - z, io.contact + 3 * m.pt_body[k], io.live, fext);
  }
  for (int k = near_ground ? HI(cyl_start, i) : 0; k < (near_ground ? HI(cyl_start, i + 1) : 0); ++k) {
    V3 c = ld3_f(m.cyl_center + 3 * k), a = ld3_f(m.cyl_axis + 3 * k);
    real rad = m.cyl_size[2 * k], hh = m.cyl_size[2 * k + 1];
    real az = dot(nrm, a);
    real s = az >= 0 ? (real)-1 : (real)1;
    V3 d = neg(nrm - az * a);
    real dn = sqrt(dot(d, d));
    V3 rim = c + (s * hh) * a;
    if (dn > (real)1e-6) rim = rim + (rad / dn) * d;
    real z = pw.z + dot(nrm, rim);
    penalty_point(p, Rw, v, rim, -z, io.contact + 3 * m.cyl_body[k], io.live, fext);
  }
  ABI I = abi_rigid(A[0], v3(A[1], A[2], A[3]), S3{A[4], A[5], A[6], A[7], A[8], A[9]});
  SV pA = crf(v, mul(I, v)) - fext;
  st6(A + A_PA, pA);
}

template <class Sync>
HD void env_substep(const EnvIO& io, real* sm, const float* hot, const DevModel& m, const SimParams& p, int g, Sync& sync) {
  const int nl = m.nl, T = m.T;
  real* X = sm + nl * LS;
  const real dt = p.dt;

  // ---- P0: joint inputs -> scratch (lane g takes links g+1, g+1+LANES, ...)
  for (int i = g; i < nl; i += DYROS_LANES) {
    real* L = sm + i * LS;
    link_inertia(io, L, hot, m, i);
    if (i == 0) continue;
    int d = HI(dof, i);
    L[LS_E] = io.dof_state[2 * d];
    L[LS_SC + 0] = io.dof_state[2 * d + 1];
    real tq = io.tau[d];
    if (p.clamp_effort) {
      real lim = HF(effort, d);
      tq = tq > lim ? lim : (tq < -lim ? -lim : tq);
    }
    L[LS_SC + 1] = tq;
    L[LS_SC + 2] = io.damping[d];
    L[LS_SC + 3] = io.armature[d];
  }
  // ---- P1: base kinematics and forces (lane 0)
  if (g == 0) {
    real* L = sm;
    V3 pw = ld3_f(io.root);
    M3 R0 = quat_to_mat(io.root[3], io.root[4], io.root[5], io.root[6]);
    SV v0{mulT(R0, ld3_f(io.root + 10)), mulT(R0, ld3_f(io.root + 7))};
    st_m3(L + LS_E, R0);
    st6(L + LS_V, v0);
    st_m3(L + LS_A + A_POSE, R0);
    st3(L + LS_A + A_POSE + 9, pw);
    link_forces(io, L, hot, m, p, 0, R0, pw, v0);
  }
  sync();
  // ---- P2: pass 1, root -> leaves: transforms, velocities, world poses, bias forces
  for (int t = 0; t < T; ++t) {
    int i = HI(sched, t * DYROS_LANES + g);
    if (i > 0) {
      real* L = sm + i * LS;
      const real* Lp = sm + HI(parent, i) * LS;
      real q = L[LS_E], qd = L[LS_SC];
      V3 ax = HF3(axis, i), r = HF3(r, i);
      M3 E = mul(axis_rot_T(ax, sin(q), cos(q)), ld_m3_f(hot + m.o_E + 9 * i));
      SV v = xform_motion(E, r, ld6(Lp + LS_V));
      v.w = v.w + qd * ax;
      M3 Rwp = ld_m3(Lp + LS_A + A_POSE);
      M3 Rw = mulABt(Rwp, E);
      V3 pw = ld3(Lp + LS_A + A_POSE + 9) + mul(Rwp, r);
      st_m3(L + LS_E, E);
      st6(L + LS_V, v);
      st_m3(L + LS_A + A_POSE, Rw);
      st3(L + LS_A + A_POSE + 9, pw);
      for (int f = 0; f < m.num_feet; ++f)
        if (m.foot_link[f] == i) {
          st_m3(X + X_FOOTPOSE + 12 * f, Rw);
          st3(X + X_FOOTPOSE + 12 * f + 9, pw);
        }
      link_forces(io, L, hot, m, p, i, Rw, pw, v);
    }
    sync();
  }
  // ---- P3: pass 2, leaves -> root: articulated inertias and bias forces
  for (int t = T - 1; t >= 0; --t) {
    int i = HI(sched, t * DYROS_LANES + g);
    if (i > 0) {
      real* L = sm + i * LS;
      real* A = L + LS_A;
      ABI IA = abi_rigid(A[0], v3(A[1], A[2], A[3]), S3{A[4], A[5], A[6], A[7], A[8], A[9]});
      SV pA = ld6(A + A_PA);
      for (int ci = HI(child_start, i); ci < HI(child_start, i + 1); ++ci) {
        const real* Ac = sm + HI(children, ci) * LS + LS_A;
        IA = IA + ld_abi(Ac + A_CIA);
        pA = pA + ld6(Ac + A_CPA);
      }
      V3 ax = HF3(axis, i), r = HF3(r, i);
      M3 E = ld_m3(L + LS_E);
      SV v = ld6(L + LS_V);
      real qd = L[LS_SC], tq = L[LS_SC + 1], damp = L[LS_SC + 2], arm = L[LS_SC + 3];
      SV U{mul(IA.I, ax), mulT(IA.H, ax)};
      real D = dot(ax, U.w) + arm + dt * damp;
      real Dinv = 1 / D;
      real u = tq - damp * qd - dot(ax, pA.w);
      V3 aq = qd * ax;
      SV c{cross(v.w, aq), cross(v.v, aq)};
      ABI Ia = rank1_sub(IA, U, Dinv);
      SV pa = pA + mul(Ia, c) + (Dinv * u) * U;
      st_abi(A + A_CIA, abi_to_parent(E, r, Ia));
      st6(A + A_CPA, xform_force_T(E, r, pa));
      st6(L + LS_U, U);
      L[LS_SC + 1] = u;
      L[LS_SC + 2] = Dinv;
      L[LS_SC + 3] = 0;
    }
    sync();
  }
  // ---- P4: floating base (lane 0): inverse articulated inertia, base acceleration, predicted base velocity
  if (g == 0) {
    real* L = sm;
    real* A = L + LS_A;
    ABI IA = abi_rigid(A[0], v3(A[1], A[2], A[3]), S3{A[4], A[5], A[6], A[7], A[8], A[9]});
    SV pA = ld6(A + A_PA);
    for (int ci = HI(child_start, 0); ci < HI(child_start, 1); ++ci) {
      const real* Ac = sm + HI(children, ci) * LS + LS_A;
      IA = IA + ld_abi(Ac + A_CIA);
      pA = pA + ld6(Ac + A_CPA);
    }
    real f[36];
    abi_to_full(IA, f);
    spd6_inverse(f);
    ABI Om0 = abi_from_full(f);
    SV a0 = (real)-1 * mul(Om0, pA);  // acceleration relative to the gravity field
    st6(A + A_ACC, a0);
    st_abi(A + A_OM0, Om0);
    M3 R0 = ld_m3(L + LS_E);
    SV v0 = ld6(L + LS_V);
    V3 gl = mulT(R0, v3(p.g[0], p.g[1], p.g[2]));
    SV vs{v0.w + dt * a0.w, v0.v + dt * (a0.v + gl)};
    st6(L + LS_U, vs);
    st3(L + LS_SC, dt * cross(v0.w, v0.v));  // rotating-frame term of the world-frame linear velocity update
  }
  sync();
  // ---- P5: pass 3, root -> leaves: joint accelerations, predicted joint velocities
  for (int t = 0; t < T; ++t) {
    int i = HI(sched, t * DYROS_LANES + g);
    if (i > 0) {
      real* L = sm + i * LS;
      const real* Lp = sm + HI(parent, i) * LS;
      V3 ax = HF3(axis, i), r = HF3(r, i);
      M3 E = ld_m3(L + LS_E);
      SV v = ld6(L + LS_V);
      real qd = L[LS_SC];
      V3 aq = qd * ax;
      SV a = xform_motion(E, r, ld6(Lp + LS_A + A_ACC)) + SV{cross(v.w, aq), cross(v.v, aq)};
      real qdd = L[LS_SC + 2] * (L[LS_SC + 1] - dot(ld6(L + LS_U), a));
      a.w = a.w + qdd * ax;
      st6(L + LS_A + A_ACC, a);
      L[LS_SC] = qd + dt * qdd;
    }
    sync();
  }
  // ---- P6: feet (lane f = foot f): inverse inertia at the foot, coupling to the base, predicted foot velocity,
  //          active sole points and their constraint rows
  const bool foot = g < m.num_feet;
  ABI Om;               // inverse inertia seen at the foot link
  SV K[6];              // rows of the map foot force -> base force
  SV V = sv_zero();     // foot velocity (foot coordinates)
  SV P = sv_zero();     // accumulated contact impulse on the foot (foot coordinates)
  M3 Rwf;
  V3 xs[MAX_ACTIVE_PTS];
  real bias[MAX_ACTIVE_PTS], lam[MAX_ACTIVE_PTS][3];
  int pbody[MAX_ACTIVE_PTS];
  int nact = 0;
  real* rows = X + X_ROWS + (foot ? g : 0) * MAX_ACTIVE_PTS * 21;
  if (foot) {
    Om = ld_abi(sm + LS_A + A_OM0);
#pragma unroll
    for (int k = 0; k < 6; ++k) K[k] = sv_zero();
    K[0].w.x = 1; K[1].w.y = 1; K[2].w.z = 1; K[3].v.x = 1; K[4].v.y = 1; K[5].v.z = 1;
    V = ld6(sm + LS_U);
    for (int k = 0; k < m.chain_len[g]; ++k) {
      int j = m.chain[g][k];
      const real* L = sm + j * LS;
      V3 ax = HF3(axis, j), r = HF3(r, j);
      M3 E = ld_m3(L + LS_E);
      real Dinv = L[LS_SC + 2];
      SV w = Dinv * xform_force_T(E, r, ld6(L + LS_U));
      SV y = mul(Om, w);
      real alpha = dot(w, y);
      Om = inv_joint_update(inv_to_child(E, r, Om), ax, xform_motion(E, r, y), alpha + Dinv);
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        real kw = dot(K[q], w);
        K[q] = xform_motion(E, r, K[q]);
        K[q].w = K[q].w - kw * ax;
      }
      V = xform_motion(E, r, V);
      V.w = V.w + L[LS_SC] * ax;
    }
    Rwf = ld_m3(X + X_FOOTPOSE + 12 * g);
    real pz = X[X_FOOTPOSE + 12 * g + 11];
    V3 nrm = v3(Rwf.a[6], Rwf.a[7], Rwf.a[8]);
#pragma unroll
    for (int a = 0; a < MAX_ACTIVE_PTS; ++a) {
      xs[a] = v3(0, 0, 0);
      bias[a] = 0;
      pbody[a] = 0;
      lam[a][0] = lam[a][1] = lam[a][2] = 0;
    }
    for (int k = 0; k < m.foot_npts[g]; ++k) {
      V3 x = v3(m.foot_pt_pos[g][k][0], m.foot_pt_pos[g][k][1], m.foot_pt_pos[g][k][2]);
      real rad = m.foot_pt_radius[g][k];
      real phi = pz + dot(nrm, x) - rad;
      if (phi < p.contact_offset && nact < MAX_ACTIVE_PTS) {
        real b = phi >= 0 ? -phi / dt : fmin_r(-p.erp * phi / dt, p.max_depen_vel);
        V3 xsk = x - rad * nrm;
#pragma unroll
        for (int a = 0; a < MAX_ACTIVE_PTS; ++a)
          if (a == nact) {
            xs[a] = xsk;
            bias[a] = b;
            pbody[a] = m.foot_pt_body[g][k];
          }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          V3 dir = d == 0 ? nrm : (d == 1 ? v3(Rwf.a[0], Rwf.a[1], Rwf.a[2]) : v3(Rwf.a[3], Rwf.a[4], Rwf.a[5]));
          SV J{cross(xsk, dir), dir};
          SV cv = mul(Om, J);
          real* rw = rows + (nact * 3 + d) * 7;
          st6(rw, cv);
          rw[6] = 1 / dot(J, cv);
        }
        ++nact;
      }
    }
  }
  // ---- P7: fixed number of sweeps; Gauss-Seidel inside a foot, Jacobi between the feet (coupled through the base)
  for (int s = 0; s < p.sweeps; ++s) {
    if (foot) {
      SV dP = sv_zero();
#pragma unroll
      for (int a = 0; a < MAX_ACTIVE_PTS; ++a) {
        if (a < nact) {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            V3 dir = d == 0 ? v3(Rwf.a[6], Rwf.a[7], Rwf.a[8])
                            : (d == 1 ? v3(Rwf.a[0], Rwf.a[1], Rwf.a[2]) : v3(Rwf.a[3], Rwf.a[4], Rwf.a[5]));
            SV J{cross(xs[a], dir), dir};
            const real* rw = rows + (a * 3 + d) * 7;
            real vrel = dot(J, V);
            real nw;
            if (d == 0) {
              nw = lam[a][0] + (bias[a] - vrel) * rw[6];
              nw = nw > 0 ? nw : 0;
            } else {
              real lim = p.mu * lam[a][0];
              nw = lam[a][d] - vrel * rw[6];
              nw = nw > lim ? lim : (nw < -lim ? -lim : nw);
            }
            real delta = nw - lam[a][d];
            lam[a][d] = nw;
            V = V + delta * ld6(rw);
            dP = dP + delta * J;
          }
        }
      }
      P = P + dP;
      SV tb{v3(dot(K[0], dP), dot(K[1], dP), dot(K[2], dP)), v3(dot(K[3], dP), dot(K[4], dP), dot(K[5], dP))};
      st6(X + X_Z + 6 * g, mul(ld_abi(sm + LS_A + A_OM0), tb));
    }
    sync();
    if (foot && m.num_feet == 2) {
      SV z = ld6(X + X_Z + 6 * (1 - g));
      V = V + z.w.x * K[0] + z.w.y * K[1] + z.w.z * K[2] + z.v.x * K[3] + z.v.y * K[4] + z.v.z * K[5];
    }
    sync();
  }
  // ---- P8: contact impulse -> joint space. Up the leg chains, base response, then down the whole tree.
  if (foot) {
    SV pd = (real)-1 * P;
    for (int k = m.chain_len[g] - 1; k >= 0; --k) {
      int j = m.chain[g][k];
      real* L = sm + j * LS;
      V3 ax = HF3(axis, j), r = HF3(r, j);
      real sd = dot(ax, pd.w);
      L[LS_SC + 3] = sd;
      pd = xform_force_T(ld_m3(L + LS_E), r, pd - (L[LS_SC + 2] * sd) * ld6(L + LS_U));
    }
    st6(X + X_PD + 6 * g, pd);
    if (io.live) {
      real inv_dt = 1 / dt;
#pragma unroll
      for (int a = 0; a < MAX_ACTIVE_PTS; ++a)
        if (a < nact) {  // world force over this sub-step: (t1, t2, n) = world (x, y, z)
          float* cf = io.contact + 3 * pbody[a];
          cf[0] += (float)(lam[a][1] * inv_dt);
          cf[1] += (float)(lam[a][2] * inv_dt);
          cf[2] += (float)(lam[a][0] * inv_dt);
        }
    }
  }
  sync();
  if (g == 0) {
    SV pd = sv_zero();
    for (int f = 0; f < m.num_feet; ++f) pd = pd + ld6(X + X_PD + 6 * f);
    st6(sm + LS_V, (real)-1 * mul(ld_abi(sm + LS_A + A_OM0), pd));
  }
  sync();
  for (int t = 0; t < T; ++t) {
    int i = HI(sched, t * DYROS_LANES + g);
    if (i > 0) {
      real* L = sm + i * LS;
      const real* Lp = sm + HI(parent, i) * LS;
      int d = HI(dof, i);
      V3 ax = HF3(axis, i), r = HF3(r, i);
      SV dv = xform_motion(ld_m3(L + LS_E), r, ld6(Lp + LS_V));
      real dqd = -L[LS_SC + 2] * (dot(ld6(L + LS_U), dv) + L[LS_SC + 3]);
      dv.w = dv.w + dqd * ax;
      st6(L + LS_V, dv);
      // joint velocity cap (dof_prop['velocity'], T:372), explicit Euler on the angle, limit projection
      real vl = HF(vel_limit, d);
      real qdn = L[LS_SC] + dqd;
      qdn = qdn > vl ? vl : (qdn < -vl ? -vl : qdn);
      real qn = io.dof_state[2 * d] + dt * qdn;
      real lo = HF(lower, d), up = HF(upper, d);
      if (qn > up) {
        qn = up;
        qdn = qdn < 0 ? qdn : 0;
      } else if (qn < lo) {
        qn = lo;
        qdn = qdn > 0 ? qdn : 0;
      }
      if (io.live) {
        io.dof_state[2 * d] = (float)qn;
        io.dof_state[2 * d + 1] = (float)qdn;
      }
    }
    sync();
  }
  // ---- P9: base integration (lane 0)
  if (g == 0) {
    const real* L = sm;
    M3 R0 = ld_m3(L + LS_E);
    SV vb = ld6(L + LS_U) + ld6(L + LS_V);
    vb.v = vb.v + ld3(L + LS_SC);
    V3 ww = mul(R0, vb.w), vw = mul(R0, vb.v);
    real wn = sqrt(dot(ww, ww));
    if (wn > p.max_ang_vel) ww = (p.max_ang_vel / wn) * ww;
    if (io.live) {
      real qx = io.root[3], qy = io.root[4], qz = io.root[5], qw = io.root[6];
      real h = (real)0.5 * dt;
      real nx = qx + h * (ww.x * qw + ww.y * qz - ww.z * qy);
      real ny = qy + h * (-ww.x * qz + ww.y * qw + ww.z * qx);
      real nz = qz + h * (ww.x * qy - ww.y * qx + ww.z * qw);
      real nw = qw + h * (-ww.x * qx - ww.y * qy - ww.z * qz);
      real inv = 1 / sqrt(nx * nx + ny * ny + nz * nz + nw * nw);
      io.root[0] = (float)(io.root[0] + dt * vw.x);
      io.root[1] = (float)(io.root[1] + dt * vw.y);
      io.root[2] = (float)(io.root[2] + dt * vw.z);
      io.root[3] = (float)(nx * inv);
      io.root[4] = (float)(ny * inv);
      io.root[5] = (float)(nz * inv);
      io.root[6] = (float)(nw * inv);
      io.root[7] = (float)vw.x; io.root[8] = (float)vw.y; io.root[9] = (float)vw.z;
      io.root[10] = (float)ww.x; io.root[11] = (float)ww.y; io.root[12] = (float)ww.z;
    }
  }
  sync();
}

}  // namespace dyros
