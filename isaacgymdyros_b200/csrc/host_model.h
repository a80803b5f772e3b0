// Host-only: DyrosModelDesc -> one packed blob of float32/int32 tables + a DevModel whose pointers are resolved
// against the blob's base address (device memory in capi.cu, host memory in tests/native/hostemu.cpp).
#pragma once
#include <string.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "internal.h"

namespace dyros {

// Host-side builder of one device blob: arrays are appended, then uploaded with a single copy.
struct Blob {
  std::vector<unsigned char> host;
  size_t add(const void* p, size_t bytes) {
    size_t off = (host.size() + 15) & ~size_t(15);
    host.resize(off + bytes);
    if (bytes) memcpy(host.data() + off, p, bytes);
    return off;
  }
  size_t add_f(const double* p, size_t n) {
    std::vector<float> f(n);
    for (size_t i = 0; i < n; ++i) f[i] = (float)p[i];
    return add(f.data(), n * sizeof(float));
  }
  size_t add_f32(const float* p, size_t n) { return add(p, n * sizeof(float)); }
  size_t add_i(const int* p, size_t n) { return add(p, n * sizeof(int)); }
};

template <class T>
static const T* at(void* base, size_t off) {
  return reinterpret_cast<const T*>(static_cast<unsigned char*>(base) + off);
}

struct ModelOffsets {
  size_t parent, dof, E, r, ax, cs, ch, bs, bd, bl, bp, br, bi, lo, up, vl, ef, ps, pb, pp, pr, ys, yb, yc, ya, yz, sc, rc, ro, dl, pg, fp, fb, st, sk;
};

// Fills `dm` (counts, foot tables) and appends every table to `bl`. Returns "" or an error message.
static std::string build_model_tables(const DyrosModelDesc* m, Blob& bl, DevModel& dm, ModelOffsets& o) {
  char err[256];
#define MFAIL(...)                        \
  do {                                    \
    snprintf(err, sizeof(err), __VA_ARGS__); \
    return std::string(err);              \
  } while (0)
  const int nl = m->num_links, nb = m->num_bodies, nd = m->num_dofs, np = m->num_points, nc = m->num_cyls;
  if (nl < 1 || nl > DYROS_MAX_LINKS) MFAIL("num_links %d outside [1,%d]", nl, DYROS_MAX_LINKS);
  if (nb < 1 || nb > DYROS_MAX_BODIES) MFAIL("num_bodies %d outside [1,%d]", nb, DYROS_MAX_BODIES);
  if (nd != nl - 1) MFAIL("one revolute DOF per non-base link expected (links %d, dofs %d)", nl, nd);
  if (m->sched_slots < 1 || !m->sched) MFAIL("missing branch schedule");
  for (int l = 1; l < nl; ++l) {
    if (m->link_parent[l] < 0 || m->link_parent[l] >= l) MFAIL("link %d parent %d not topological", l, m->link_parent[l]);
    if (m->link_dof[l] < 0 || m->link_dof[l] >= nd) MFAIL("link %d dof index %d", l, m->link_dof[l]);
  }
  // `sched` holds the role programs (model/tables.py::role_programs): column r = links of role r in ascending
  // (topological) order, every link exactly once.
  std::vector<int> role_of(nl, -1);
  int role_len[DYROS_LANES] = {0};
  for (int g = 0; g < DYROS_LANES; ++g) {
    int prev = 0;
    for (int t = 0; t < m->sched_slots; ++t) {
      int l = m->sched[t * DYROS_LANES + g];
      if (l < 0) continue;
      if (l < 1 || l >= nl || role_of[l] >= 0) MFAIL("bad role program entry %d", l);
      if (l <= prev || role_len[g] != t) MFAIL("role %d is not a dense ascending list at row %d", g, t);
      prev = l;
      role_of[l] = g;
      role_len[g] = t + 1;
    }
  }
  for (int l = 1; l < nl; ++l)
    if (role_of[l] < 0) MFAIL("link %d missing from the role programs", l);
  // ---- derived tables
  std::vector<int> child_start(nl + 1, 0), children;
  for (int l = 0; l < nl; ++l) {
    child_start[l] = (int)children.size();
    for (int c = 1; c < nl; ++c)
      if (m->link_parent[c] == l) children.push_back(c);
  }
  child_start[nl] = (int)children.size();
  std::vector<int> body_start(nl + 1, 0), bodies;
  for (int l = 0; l < nl; ++l) {
    body_start[l] = (int)bodies.size();
    for (int bb = 0; bb < nb; ++bb)
      if (m->body_link[bb] == l) bodies.push_back(bb);
  }
  body_start[nl] = (int)bodies.size();
  if ((int)bodies.size() != nb) MFAIL("body_link has entries outside [0,%d)", nl);
  std::vector<int> pt_start(nl + 1, 0), ppt_body;
  std::vector<float> ppt_pos, ppt_rad;
  for (int l = 0; l < nl; ++l) {
    pt_start[l] = (int)ppt_body.size();
    for (int i = 0; i < np; ++i)
      if (m->pt_link[i] == l && !(m->pt_solver && m->pt_solver[i])) {
        ppt_body.push_back(m->pt_body[i]);
        for (int k = 0; k < 3; ++k) ppt_pos.push_back((float)m->pt_pos[3 * i + k]);
        ppt_rad.push_back((float)m->pt_radius[i]);
      }
  }
  pt_start[nl] = (int)ppt_body.size();
  std::vector<int> cyl_start(nl + 1, 0), ccyl_body;
  std::vector<float> ccyl_center, ccyl_axis, ccyl_size;
  for (int l = 0; l < nl; ++l) {
    cyl_start[l] = (int)ccyl_body.size();
    for (int i = 0; i < nc; ++i)
      if (m->cyl_link[i] == l) {
        ccyl_body.push_back(m->cyl_body[i]);
        for (int k = 0; k < 3; ++k) ccyl_center.push_back((float)m->cyl_center[3 * i + k]);
        for (int k = 0; k < 3; ++k) ccyl_axis.push_back((float)m->cyl_axis[3 * i + k]);
        for (int k = 0; k < 2; ++k) ccyl_size.push_back((float)m->cyl_size[2 * i + k]);
      }
  }
  cyl_start[nl] = (int)ccyl_body.size();

  memset(&dm, 0, sizeof(dm));
  dm.nl = nl; dm.nb = nb; dm.nd = nd; dm.np = (int)ppt_body.size(); dm.nc = (int)ccyl_body.size(); dm.T = m->sched_slots;
  // solver (foot) links, their chains and candidate points, in ascending link order
  for (int i = 0; i < np; ++i) {
    if (!(m->pt_solver && m->pt_solver[i])) continue;
    int l = m->pt_link[i], f = -1;
    for (int k = 0; k < dm.num_feet; ++k)
      if (dm.foot_link[k] == l) f = k;
    if (f < 0) {
      if (dm.num_feet >= MAX_FEET) MFAIL("more than %d solver links", MAX_FEET);
      f = dm.num_feet++;
      dm.foot_link[f] = l;
    }
    if (dm.foot_npts[f] >= MAX_SOLVER_PTS) MFAIL("more than %d solver points on link %d", MAX_SOLVER_PTS, l);
    int k = dm.foot_npts[f]++;
    dm.foot_pt_body[f][k] = m->pt_body[i];
    for (int c = 0; c < 3; ++c) dm.foot_pt_pos[f][k][c] = (float)m->pt_pos[3 * i + c];
    dm.foot_pt_radius[f][k] = (float)m->pt_radius[i];
  }
  if (dm.num_feet == 2 && dm.foot_link[0] > dm.foot_link[1]) {
    std::swap(dm.foot_link[0], dm.foot_link[1]);
    std::swap(dm.foot_npts[0], dm.foot_npts[1]);
    for (int k = 0; k < MAX_SOLVER_PTS; ++k) {
      std::swap(dm.foot_pt_body[0][k], dm.foot_pt_body[1][k]);
      std::swap(dm.foot_pt_radius[0][k], dm.foot_pt_radius[1][k]);
      for (int c = 0; c < 3; ++c) std::swap(dm.foot_pt_pos[0][k][c], dm.foot_pt_pos[1][k][c]);
    }
  }
  // Leg chains hang off the lowest common ancestor (LCA) of the solver links (the base for TOCABI, the pelvis for the
  // Humanoid); `shared` = links from the base down to the LCA (exclusive of the base, inclusive of the LCA).
  {
    std::vector<std::vector<int>> paths(dm.num_feet);
    for (int f = 0; f < dm.num_feet; ++f) {
      for (int l = dm.foot_link[f]; l > 0; l = m->link_parent[l]) paths[f].push_back(l);
      std::reverse(paths[f].begin(), paths[f].end());
    }
    size_t common = 0;
    if (dm.num_feet == 2)
      while (common < paths[0].size() && common < paths[1].size() && paths[0][common] == paths[1][common]) ++common;
    if (common > (size_t)MAX_CHAIN) MFAIL("%zu links between the base and the feet's common ancestor (max %d)", common, MAX_CHAIN);
    dm.shared_len = (int)common;
    for (size_t k = 0; k < common; ++k) dm.shared[k] = paths[0][k];
    dm.lca = common ? paths[0][common - 1] : 0;
    for (int f = 0; f < dm.num_feet; ++f) {
      const int len = (int)paths[f].size() - (int)common;
      if (len > MAX_CHAIN || len < 1) MFAIL("solver link %d is %d joints below the common ancestor (1..%d)", dm.foot_link[f], len, MAX_CHAIN);
      if (len * 2 < MAX_ACTIVE_PTS * 3)
        MFAIL("solver link %d is only %d joints below the common ancestor; the contact rows are parked in the chain's "
              "scratch blocks (2 per link, %d needed)", dm.foot_link[f], len, MAX_ACTIVE_PTS * 3);
      dm.chain_len[f] = len;
      for (int k = 0; k < len; ++k) dm.chain[f][k] = paths[f][common + k];
    }
  }
  for (int f = 0; f < dm.num_feet; ++f) {
    dm.foot_role[f] = role_of[dm.foot_link[f]];
    for (int k = 0; k < dm.chain_len[f]; ++k)
      if (role_of[dm.chain[f][k]] != dm.foot_role[f]) MFAIL("leg chain of solver link %d is split between roles", dm.foot_link[f]);
    for (int f2 = 0; f2 < f; ++f2)
      if (dm.foot_role[f2] == dm.foot_role[f]) MFAIL("two solver links share role %d", dm.foot_role[f]);
  }
  {  // the base is handled by the least loaded role that has no foot
    int best = -1;
    for (int g = 0; g < DYROS_LANES; ++g) {
      bool is_foot = false;
      for (int f = 0; f < dm.num_feet; ++f) is_foot |= dm.foot_role[f] == g;
      if (!is_foot && (best < 0 || role_len[g] < role_len[best])) best = g;
    }
    if (best < 0) best = 0;
    dm.base_role = best;
    role_of[0] = best;
    for (int g = 0; g < DYROS_LANES; ++g) dm.role_len[g] = role_len[g];
  }
  // per-link bounding radius of the penalty candidates about the link origin (early-out of the contact loops)
  std::vector<float> reach(nl, 0.f);
  for (int l = 0; l < nl; ++l) {
    for (int k = pt_start[l]; k < pt_start[l + 1]; ++k) {
      float x = ppt_pos[3 * k], y = ppt_pos[3 * k + 1], z = ppt_pos[3 * k + 2];
      reach[l] = std::max(reach[l], std::sqrt(x * x + y * y + z * z) + ppt_rad[k]);
    }
    for (int k = cyl_start[l]; k < cyl_start[l + 1]; ++k) {
      float x = ccyl_center[3 * k], y = ccyl_center[3 * k + 1], z = ccyl_center[3 * k + 2];
      reach[l] = std::max(reach[l], std::sqrt(x * x + y * y + z * z) + ccyl_size[2 * k] + ccyl_size[2 * k + 1]);
    }
  }
  // hot tables first: the physics kernel stages the prefix [0, hot_bytes) into shared memory
  o.bi = bl.add_f(m->body_inertia, nb * 10);
  o.dof = bl.add_i(m->link_dof, nl);
  {  // flattened role programs
    std::vector<int> order(1, 0);
    for (int g = 0; g < DYROS_LANES; ++g) {
      dm.prog_start[g] = (int)order.size();
      for (int t = 0; t < role_len[g]; ++t) order.push_back(m->sched[t * DYROS_LANES + g]);
    }
    std::vector<int> rec_of(nl, 0);
    for (size_t k = 0; k < order.size(); ++k) rec_of[order[k]] = (int)k;
    std::vector<int> prog(order.size() * REC_WORDS, 0);
    int n_cslot = 0;
    auto F = [&](size_t k, int field) -> float& { return reinterpret_cast<float*>(prog.data())[k * REC_WORDS + field]; };
    for (size_t k = 0; k < order.size(); ++k) {
      const int l = order[k];
      int* R = prog.data() + k * REC_WORDS;
      const int par = l > 0 ? m->link_parent[l] : 0;
      R[R_LINK] = l;
      R[R_PARENT] = par;
      R[R_FLAGS] = (l > 0 && role_of[par] != role_of[l] ? RF_PARENT_FOREIGN : 0) | (l > 0 && par == 0 ? RF_PARENT_BASE : 0);
      for (int ci = child_start[l]; ci < child_start[l + 1]; ++ci)
        if (role_of[children[ci]] != role_of[l]) R[R_FLAGS] |= RF_PUBLISH;
      R[R_DOF] = l > 0 ? m->link_dof[l] : 0;
      const int nbod = body_start[l + 1] - body_start[l], nch = child_start[l + 1] - child_start[l];
      if (nbod > MAX_LINK_BODIES) MFAIL("link %d merges %d bodies (max %d)", l, nbod, MAX_LINK_BODIES);
      if (nch > MAX_LINK_CHILDREN) MFAIL("link %d has %d children (max %d)", l, nch, MAX_LINK_CHILDREN);
      R[R_NBODY] = nbod;
      for (int j = 0; j < nbod; ++j) R[R_BODY0 + j] = bodies[body_start[l] + j];
      R[R_NCHILD] = nch;
      for (int j = 0; j < nch; ++j) {
        int c = children[child_start[l] + j];
        R[R_CHILD0 + j] = c | (role_of[c] != role_of[l] ? REC_FOREIGN : 0);
      }
      for (int j = 0; j < 3; ++j) F(k, R_AXIS + j) = (float)m->link_axis[3 * l + j];
      for (int j = 0; j < 3; ++j) F(k, R_R + j) = (float)m->link_r[3 * l + j];
      for (int j = 0; j < 9; ++j) F(k, R_E + j) = (float)m->link_E[9 * l + j];
      F(k, R_REACH) = reach[l];
      R[R_PT0] = pt_start[l]; R[R_PT1] = pt_start[l + 1];
      R[R_CYL0] = cyl_start[l]; R[R_CYL1] = cyl_start[l + 1];
      if (l > 0) {
        const int d = m->link_dof[l];
        F(k, R_VLIM) = (float)m->dof_vel_limit[d];
        F(k, R_LO) = (float)m->dof_lower[d];
        F(k, R_UP) = (float)m->dof_upper[d];
        F(k, R_EFF) = (float)m->dof_effort[d];
        F(k, R_STIFF) = m->dof_stiffness ? (float)m->dof_stiffness[d] : 0.f;
      }
      R[R_FOOT] = -1;
      for (int f = 0; f < dm.num_feet; ++f)
        if (dm.foot_link[f] == l) R[R_FOOT] = f;
      // multi-lane program: results travel in registers between consecutive links of a role; everything else goes
      // through the scratch block. R_CSLOT: this link's parent is not the previous link of the role (the base, for the
      // first link of the base role); RF_KEEP: a same-role child of this link is not the role's next link.
      R[R_CSLOT] = -1;
      if (l > 0) {
        const int g = role_of[l];
        int t = 0;
        while (m->sched[t * DYROS_LANES + g] != l) ++t;
        const int before = t > 0 ? m->sched[(t - 1) * DYROS_LANES + g] : (g == dm.base_role ? 0 : -1);
        if (par != before) {
          if (n_cslot >= MAX_CSLOTS) MFAIL("more than %d links whose parent is not their role's previous link", MAX_CSLOTS);
          R[R_CSLOT] = n_cslot++;
        }
        const int after = t + 1 < role_len[g] ? m->sched[(t + 1) * DYROS_LANES + g] : -1;
        for (int ci = child_start[l]; ci < child_start[l + 1]; ++ci)
          if (role_of[children[ci]] == g && children[ci] != after) R[R_FLAGS] |= RF_KEEP;
      }
    }
    o.pg = bl.add_i(prog.data(), prog.size());
    for (int f = 0; f < dm.num_feet; ++f)
      for (int k = 0; k < dm.chain_len[f]; ++k) dm.chain_rec[f][k] = rec_of[dm.chain[f][k]];
    for (int k = 0; k < dm.shared_len; ++k) dm.shared_rec[k] = rec_of[dm.shared[k]];
    for (int g = 0; g < DYROS_LANES; ++g) {
      dm.n_xchild[g] = 0;
      for (int t = 0; t < role_len[g]; ++t) {
        int l = m->sched[t * DYROS_LANES + g];
        for (int ci = child_start[l]; ci < child_start[l + 1]; ++ci)
          if (role_of[children[ci]] != g) {
            if (dm.n_xchild[g] >= 8) MFAIL("role %d has more than 8 children in other roles", g);
            dm.xchild[g][dm.n_xchild[g]++] = children[ci];
          }
      }
    }
  }
  std::vector<int> dof_link(nd, 0);
  for (int l = 1; l < nl; ++l) dof_link[m->link_dof[l]] = l;
  o.dl = bl.add_i(dof_link.data(), nd);
  {  // solver-point candidates of the feet (read with a per-lane index by the multi-lane kernel: staged, not in DevModel)
    std::vector<float> fp(MAX_FEET * MAX_SOLVER_PTS * 4, 0.f);
    std::vector<int> fb(MAX_FEET * MAX_SOLVER_PTS, 0);
    for (int f = 0; f < dm.num_feet; ++f)
      for (int k = 0; k < dm.foot_npts[f]; ++k) {
        for (int c = 0; c < 3; ++c) fp[(f * MAX_SOLVER_PTS + k) * 4 + c] = dm.foot_pt_pos[f][k][c];
        fp[(f * MAX_SOLVER_PTS + k) * 4 + 3] = dm.foot_pt_radius[f][k];
        fb[f * MAX_SOLVER_PTS + k] = dm.foot_pt_body[f][k];
      }
    o.fp = bl.add_f32(fp.data(), fp.size());
    o.fb = bl.add_i(fb.data(), fb.size());
  }
  dm.hot_bytes = (int)((bl.host.size() + 15) & ~size_t(15));
  // cold tables (global memory): penalty candidates (read only when a link is near the ground), rigid_body_state
  // kinematics, and the unflattened tree
  o.parent = bl.add_i(m->link_parent, nl);
  o.E = bl.add_f(m->link_E, nl * 9); o.r = bl.add_f(m->link_r, nl * 3); o.ax = bl.add_f(m->link_axis, nl * 3);
  o.cs = bl.add_i(child_start.data(), nl + 1); o.ch = bl.add_i(children.data(), children.size());
  o.bs = bl.add_i(body_start.data(), nl + 1); o.bd = bl.add_i(bodies.data(), nb);
  o.lo = bl.add_f(m->dof_lower, nd); o.up = bl.add_f(m->dof_upper, nd); o.vl = bl.add_f(m->dof_vel_limit, nd);
  o.ef = bl.add_f(m->dof_effort, nd);
  {
    std::vector<double> st(nd, 0.0);
    if (m->dof_stiffness)
      for (int d = 0; d < nd; ++d) st[d] = m->dof_stiffness[d];
    o.st = bl.add_f(st.data(), nd);
  }
  o.ps = bl.add_i(pt_start.data(), nl + 1); o.ys = bl.add_i(cyl_start.data(), nl + 1);
  o.sc = bl.add_i(m->sched, (size_t)m->sched_slots * DYROS_LANES);
  o.rc = bl.add_f32(reach.data(), nl);
  o.ro = bl.add_i(role_of.data(), nl);
  o.bl = bl.add_i(m->body_link, nb); o.bp = bl.add_f(m->body_pos, nb * 3); o.br = bl.add_f(m->body_rot, nb * 9);
  o.pb = bl.add_i(ppt_body.data(), ppt_body.size());
  o.pp = bl.add_f32(ppt_pos.data(), ppt_pos.size()); o.pr = bl.add_f32(ppt_rad.data(), ppt_rad.size());
  o.yb = bl.add_i(ccyl_body.data(), ccyl_body.size());
  o.yc = bl.add_f32(ccyl_center.data(), ccyl_center.size()); o.ya = bl.add_f32(ccyl_axis.data(), ccyl_axis.size());
  o.yz = bl.add_f32(ccyl_size.data(), ccyl_size.size());
  {  // self-collision tables: one packed block (layout in internal.h)
    const int ns = m->sc_shape_kind ? m->sc_num_shapes : 0, nsm = ns ? m->sc_num_samples : 0, npair = ns ? m->sc_num_pairs : 0;
    dm.sc_ns = ns; dm.sc_nsamp = nsm; dm.sc_np = npair;
    if (ns > SC_MAX_SHAPES) MFAIL("self-collision: %d shapes, at most %d", ns, SC_MAX_SHAPES);
    if (ns && (nl > 255 || nb > 255 || nsm > 65535)) MFAIL("self-collision: the model is too large for the packed tables");
    std::vector<int> hot;
    auto section = [&]() { hot.resize((hot.size() + 3) & ~size_t(3)); return (int)hot.size(); };
    auto f2i = [](float f) { int i; memcpy(&i, &f, 4); return i; };
    std::vector<unsigned short> sp;
    std::vector<int> q0(npair + 1, 0);
    for (int k = 0; k < npair; ++k) {
      const int li = m->sc_pairs[2 * k], lj = m->sc_pairs[2 * k + 1];
      if (li < 0 || li >= nl || lj < 0 || lj >= nl) MFAIL("self-collision pair %d: link out of range", k);
      for (int a = m->sc_link_shape0[li]; a < m->sc_link_shape0[li + 1]; ++a)
        for (int b = m->sc_link_shape0[lj]; b < m->sc_link_shape0[lj + 1]; ++b) sp.push_back((unsigned short)(a | b << 8));
      q0[k + 1] = (int)sp.size();
    }
    dm.sc_nq = (int)sp.size();
    // padded to whole batches of the kernel's sweep with a pair of two far-apart dummy shapes (ns, ns + 1); at most 16
    // chunks (the kernel keeps one bit per chunk), each a whole number of batches
    while (sp.size() % SC_SWEEP_BATCH) sp.push_back((unsigned short)(ns | (ns + 1) << 8));
    {
      const int nbatch = (int)sp.size() / SC_SWEEP_BATCH;
      dm.sc_chunk = SC_SWEEP_BATCH * std::max(1, (nbatch + 15) / 16);
      while (sp.size() % dm.sc_chunk) sp.push_back((unsigned short)(ns | (ns + 1) << 8));
    }
    dm.sc_nq_padded = (int)sp.size();
    dm.sc_o_pair = section();
    for (int k = 0; k < npair; ++k) {
      unsigned mask = 0;
      for (int c = q0[k] / dm.sc_chunk; c <= (q0[k + 1] - 1) / dm.sc_chunk && q0[k + 1] > q0[k]; ++c) mask |= 1u << c;
      hot.push_back((int)((unsigned)m->sc_pairs[2 * k] | (unsigned)m->sc_pairs[2 * k + 1] << 8 | mask << 16));
    }
    dm.sc_o_lsph = section();
    for (int k = 0; k < (ns ? nl * 4 : 0); ++k) hot.push_back(f2i((float)m->sc_link_sphere[k]));
    dm.sc_o_sp = section();
    for (size_t k = 0; k < sp.size(); k += 2) hot.push_back((int)(sp[k] | (unsigned)sp[k + 1] << 16));
    dm.sc_o_shape = section();
    for (int k = 0; k < ns; ++k) {
      if (m->sc_shape_link[k] < 0 || m->sc_shape_link[k] >= nl || m->sc_shape_body[k] < 0 || m->sc_shape_body[k] >= nb)
        MFAIL("self-collision shape %d: link / body out of range", k);
      for (int c = 0; c < 3; ++c) hot.push_back(f2i((float)m->sc_shape_center[3 * k + c]));
      for (int c = 0; c < 9; ++c) hot.push_back(f2i((float)m->sc_shape_rot[9 * k + c]));
      for (int c = 0; c < 3; ++c) hot.push_back(f2i((float)m->sc_shape_size[3 * k + c]));
      // bounding radius about the centre (rounded up): box |half extents|, cylinder hypot(radius, half height),
      // capsule radius + half height
      const double* z = m->sc_shape_size + 3 * k;
      if (m->sc_shape_kind[k] < 0 || m->sc_shape_kind[k] > 2) MFAIL("self-collision shape %d: kind %d", k, m->sc_shape_kind[k]);
      const double R = m->sc_shape_kind[k] == 0 ? std::sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2])
                       : (m->sc_shape_kind[k] == 1 ? std::hypot(z[0], z[1]) : z[0] + z[1]);
      hot.push_back(f2i(std::nextafter((float)(R * (1.0 + 1e-6)), INFINITY)));
    }
    dm.sc_o_meta = section();
    for (int k = 0; k < ns; ++k) {
      const int s0 = m->sc_shape_sample0[k], n = m->sc_shape_sample0[k + 1] - s0;
      if (n < 0 || n > SC_MAX_SHAPE_SAMPLES || s0 < 0 || s0 + n > nsm)
        MFAIL("self-collision shape %d: %d sample spheres (at most %d)", k, n, SC_MAX_SHAPE_SAMPLES);
      hot.push_back(m->sc_shape_link[k] | m->sc_shape_body[k] << 8 | m->sc_shape_kind[k] << 16);
      hot.push_back(s0 | n << 16);
    }
    dm.sc_o_sample = section();
    for (int k = 0; k < nsm * 4; ++k) hot.push_back(f2i((float)m->sc_sample[k]));
    section();
    if (hot.empty()) hot.resize(4, 0);
    dm.sc_hot_words = (int)hot.size();
    o.sk = bl.add_i(hot.data(), hot.size());
  }
  bl.host.resize((bl.host.size() + 15) & ~size_t(15));
  // word offsets of the hot tables inside the staged prefix
  dm.o_parent = (int)(o.parent / 4); dm.o_dof = (int)(o.dof / 4); dm.o_E = (int)(o.E / 4); dm.o_r = (int)(o.r / 4);
  dm.o_axis = (int)(o.ax / 4); dm.o_child_start = (int)(o.cs / 4); dm.o_children = (int)(o.ch / 4);
  dm.o_body_start = (int)(o.bs / 4); dm.o_bodies = (int)(o.bd / 4); dm.o_body_inertia = (int)(o.bi / 4);
  dm.o_lower = (int)(o.lo / 4); dm.o_upper = (int)(o.up / 4); dm.o_vel_limit = (int)(o.vl / 4);
  dm.o_effort = (int)(o.ef / 4); dm.o_pt_start = (int)(o.ps / 4); dm.o_cyl_start = (int)(o.ys / 4);
  dm.o_sched = (int)(o.sc / 4); dm.o_reach = (int)(o.rc / 4); dm.o_role_of = (int)(o.ro / 4); dm.o_dof_link = (int)(o.dl / 4); dm.o_prog = (int)(o.pg / 4);
  dm.o_foot_pts = (int)(o.fp / 4); dm.o_foot_body = (int)(o.fb / 4);

  return std::string();
#undef MFAIL
}

static void resolve_model(DevModel& dm, const ModelOffsets& o, void* base) {
  dm.link_parent = at<int>(base, o.parent); dm.link_dof = at<int>(base, o.dof);
  dm.link_E = at<float>(base, o.E); dm.link_r = at<float>(base, o.r); dm.link_axis = at<float>(base, o.ax);
  dm.link_child_start = at<int>(base, o.cs); dm.link_children = at<int>(base, o.ch);
  dm.link_body_start = at<int>(base, o.bs); dm.link_bodies = at<int>(base, o.bd);
  dm.body_link = at<int>(base, o.bl); dm.body_pos = at<float>(base, o.bp); dm.body_rot = at<float>(base, o.br);
  dm.body_inertia = at<float>(base, o.bi);
  dm.dof_lower = at<float>(base, o.lo); dm.dof_upper = at<float>(base, o.up); dm.dof_vel_limit = at<float>(base, o.vl);
  dm.dof_effort = at<float>(base, o.ef);
  dm.dof_stiffness = at<float>(base, o.st);
  dm.link_pt_start = at<int>(base, o.ps); dm.pt_body = at<int>(base, o.pb); dm.pt_pos = at<float>(base, o.pp);
  dm.pt_radius = at<float>(base, o.pr);
  dm.link_cyl_start = at<int>(base, o.ys); dm.cyl_body = at<int>(base, o.yb); dm.cyl_center = at<float>(base, o.yc);
  dm.cyl_axis = at<float>(base, o.ya); dm.cyl_size = at<float>(base, o.yz);
  dm.sched = at<int>(base, o.sc);
  dm.sc_hot = at<int>(base, o.sk);
  dm.link_reach = at<float>(base, o.rc);
  dm.blob = base;
}

static void fill_sim_params(const DyrosSimDesc* d, SimParams& p) {
  p.N = d->num_envs;
  p.substeps = d->substeps;
  p.dt = (float)(d->dt / (double)d->substeps);
  for (int i = 0; i < 3; ++i) p.g[i] = d->gravity[i];
  p.contact_offset = d->contact_offset;
  p.max_depen_vel = d->max_depenetration_velocity;
  p.erp = d->contact_erp;
  p.mu = d->friction;
  p.pen_k = d->penalty_stiffness;
  p.pen_c = d->penalty_damping;
  p.pen_fmax = d->penalty_max_force;
  p.max_ang_vel = d->max_angular_velocity;
  p.sweeps = d->contact_sweeps;
  p.clamp_effort = d->clamp_effort;
  p.sc_hits_cap = 0;  // set at the first launch_self_collision
}

}  // namespace dyros
