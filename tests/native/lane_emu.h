// TEST INFRASTRUCTURE ONLY. Host emulation of one lane group of the multi-lane physics program
// (isaacgymdyros_b200/csrc/lanes.cuh describes the vocabulary): every `real` is LPE scalars, one per lane; shuffles
// are permutations; replicated stores check that all lanes agree (a lane-varying value reaching a replicated store is a
// bug in the program). EMU_SCALAR = double gives a float64 run of the same program (separates rounding from
// formulation errors). Include this BEFORE any header of the program.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#define DYROS_LANE_EMU 1
#ifndef EMU_SCALAR
#define EMU_SCALAR float
#endif

constexpr int LPE = 8;
typedef EMU_SCALAR emu_s;

struct LaneMask {
  bool v[LPE];
};
struct LaneInt {
  int v[LPE];
  LaneInt() {}
  LaneInt(int x) {
    for (int i = 0; i < LPE; ++i) v[i] = x;
  }
};
struct LaneVec {
  emu_s v[LPE];
  LaneVec() {}
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  LaneVec(T x) {
    for (int i = 0; i < LPE; ++i) v[i] = (emu_s)x;
  }
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  explicit operator T() const {  // only meaningful for replicated values
    return (T)v[0];
  }
};
#define LV_BIN(op)                                                   \
  inline LaneVec operator op(const LaneVec& a, const LaneVec& b) {   \
    LaneVec r;                                                       \
    for (int i = 0; i < LPE; ++i) r.v[i] = a.v[i] op b.v[i];         \
    return r;                                                        \
  }                                                                  \
  inline LaneVec& operator op##=(LaneVec& a, const LaneVec& b) {     \
    for (int i = 0; i < LPE; ++i) a.v[i] = a.v[i] op b.v[i];         \
    return a;                                                        \
  }
LV_BIN(+) LV_BIN(-) LV_BIN(*) LV_BIN(/)
inline LaneVec operator-(const LaneVec& a) {
  LaneVec r;
  for (int i = 0; i < LPE; ++i) r.v[i] = -a.v[i];
  return r;
}
#define LV_CMP(op)                                                   \
  inline LaneMask operator op(const LaneVec& a, const LaneVec& b) {  \
    LaneMask r;                                                      \
    for (int i = 0; i < LPE; ++i) r.v[i] = a.v[i] op b.v[i];         \
    return r;                                                        \
  }
LV_CMP(<) LV_CMP(>) LV_CMP(<=) LV_CMP(>=) LV_CMP(==)
#define LI_BIN(op)                                                   \
  inline LaneInt operator op(const LaneInt& a, const LaneInt& b) {   \
    LaneInt r;                                                       \
    for (int i = 0; i < LPE; ++i) r.v[i] = a.v[i] op b.v[i];         \
    return r;                                                        \
  }
LI_BIN(+) LI_BIN(-) LI_BIN(*) LI_BIN(/) LI_BIN(%) LI_BIN(&)
#define LI_CMP(op)                                                   \
  inline LaneMask operator op(const LaneInt& a, const LaneInt& b) {  \
    LaneMask r;                                                      \
    for (int i = 0; i < LPE; ++i) r.v[i] = a.v[i] op b.v[i];         \
    return r;                                                        \
  }
LI_CMP(<) LI_CMP(>) LI_CMP(<=) LI_CMP(>=) LI_CMP(==) LI_CMP(!=)
inline LaneMask operator&&(const LaneMask& a, const LaneMask& b) {
  LaneMask r;
  for (int i = 0; i < LPE; ++i) r.v[i] = a.v[i] && b.v[i];
  return r;
}
inline LaneMask operator||(const LaneMask& a, const LaneMask& b) {
  LaneMask r;
  for (int i = 0; i < LPE; ++i) r.v[i] = a.v[i] || b.v[i];
  return r;
}
inline LaneMask operator!(const LaneMask& a) {
  LaneMask r;
  for (int i = 0; i < LPE; ++i) r.v[i] = !a.v[i];
  return r;
}
inline LaneMask operator&&(const LaneMask& a, bool b) {
  LaneMask r;
  for (int i = 0; i < LPE; ++i) r.v[i] = a.v[i] && b;
  return r;
}
inline LaneMask operator&&(bool b, const LaneMask& a) { return a && b; }
inline LaneVec sqrt(const LaneVec& a) {
  LaneVec r;
  for (int i = 0; i < LPE; ++i) r.v[i] = (emu_s)::sqrt((double)a.v[i]);
  return r;
}
inline LaneVec sin(const LaneVec& a) {
  LaneVec r;
  for (int i = 0; i < LPE; ++i) r.v[i] = (emu_s)::sin((double)a.v[i]);
  return r;
}
inline LaneVec cos(const LaneVec& a) {
  LaneVec r;
  for (int i = 0; i < LPE; ++i) r.v[i] = (emu_s)::cos((double)a.v[i]);
  return r;
}

#define DYROS_REAL LaneVec

namespace dyros {
typedef LaneInt li;
typedef LaneMask lb;

struct Ln {
  int dummy;
};
inline Ln make_ln(int) { return Ln{0}; }
inline li lane_index(const Ln&) {
  li r;
  for (int i = 0; i < LPE; ++i) r.v[i] = i;
  return r;
}
[[noreturn]] inline void emu_die(const char* what) {
  fprintf(stderr, "lane emulation: %s\n", what);
  abort();
}
inline LaneVec bc(const Ln&, const LaneVec& x, int k) { return LaneVec(x.v[k]); }
inline LaneVec sh(const Ln&, const LaneVec& x, const LaneInt& src) {
  LaneVec r;
  for (int i = 0; i < LPE; ++i) {
    if (src.v[i] < 0 || src.v[i] >= LPE) emu_die("shuffle source outside the group");
    r.v[i] = x.v[src.v[i]];
  }
  return r;
}
inline li grank(const Ln&, const LaneMask& m) {
  li r;
  int n = 0;
  for (int i = 0; i < LPE; ++i) {
    r.v[i] = n;
    n += m.v[i] ? 1 : 0;
  }
  return r;
}
inline int gcount(const Ln&, const LaneMask& m) {
  int n = 0;
  for (int i = 0; i < LPE; ++i) n += m.v[i] ? 1 : 0;
  return n;
}
inline LaneVec ld(const float* p) { return LaneVec(*p); }
inline void check_replicated(const LaneVec& x, const char* what) {
  for (int i = 1; i < LPE; ++i)
    if (!(x.v[i] == x.v[0]) && !(x.v[i] != x.v[i] && x.v[0] != x.v[0])) emu_die(what);
}
inline void st(float* p, const LaneVec& x) {
  check_replicated(x, "replicated store of a lane-varying value");
  *p = (float)x.v[0];
}
inline int ldi(const int* p) { return *p; }
inline void ld4(const float* p, LaneVec& a, LaneVec& b, LaneVec& c, LaneVec& d) {
  if (reinterpret_cast<uintptr_t>(p) & 15) emu_die("ld4 of an address that is not 16-byte aligned");
  a = LaneVec(p[0]); b = LaneVec(p[1]); c = LaneVec(p[2]); d = LaneVec(p[3]);
}
inline void st4(float* p, const LaneVec& a, const LaneVec& b, const LaneVec& c, const LaneVec& d) {
  if (reinterpret_cast<uintptr_t>(p) & 15) emu_die("st4 of an address that is not 16-byte aligned");
  st(p, a); st(p + 1, b); st(p + 2, c); st(p + 3, d);
}
inline LaneVec ldl(const float* p, const LaneInt& i) {
  LaneVec r;
  for (int k = 0; k < LPE; ++k) r.v[k] = p[i.v[k]];
  return r;
}
inline LaneInt ldli(const int* p, const LaneInt& i) {
  LaneInt r;
  for (int k = 0; k < LPE; ++k) r.v[k] = p[i.v[k]];
  return r;
}
inline void stl(float* p, const LaneInt& i, const LaneVec& x, const LaneMask& m) {
  for (int k = 0; k < LPE; ++k)
    if (m.v[k]) p[i.v[k]] = (float)x.v[k];
}
inline void stli(int* p, const LaneInt& i, const LaneInt& x, const LaneMask& m) {
  for (int k = 0; k < LPE; ++k)
    if (m.v[k]) p[i.v[k]] = x.v[k];
}
inline LaneVec sel(const LaneMask& m, const LaneVec& a, const LaneVec& b) {
  LaneVec r;
  for (int i = 0; i < LPE; ++i) r.v[i] = m.v[i] ? a.v[i] : b.v[i];
  return r;
}
inline LaneInt seli(const LaneMask& m, const LaneInt& a, const LaneInt& b) {
  LaneInt r;
  for (int i = 0; i < LPE; ++i) r.v[i] = m.v[i] ? a.v[i] : b.v[i];
  return r;
}
inline bool any(const LaneMask& m) {
  for (int i = 0; i < LPE; ++i)
    if (m.v[i]) return true;
  return false;
}
inline bool uni(const LaneMask& m) {
  for (int i = 1; i < LPE; ++i)
    if (m.v[i] != m.v[0]) emu_die("uni() of a predicate that differs between the lanes");
  return m.v[0];
}
inline int unii(const LaneInt& x) {
  for (int i = 1; i < LPE; ++i)
    if (x.v[i] != x.v[0]) emu_die("unii() of an int that differs between the lanes");
  return x.v[0];
}
inline LaneVec rcp_r(const LaneVec& x) { return LaneVec(1) / x; }
inline LaneVec sqrt_r(const LaneVec& x) { return sqrt(x); }
inline LaneVec int_bits_as_real(const LaneInt& x) {
  LaneVec r;
  for (int i = 0; i < LPE; ++i) {
    float f;
    memcpy(&f, &x.v[i], 4);
    r.v[i] = f;  // (exact: a float value widened to the emulation scalar)
  }
  return r;
}
inline LaneInt real_bits_as_int(const LaneVec& x) {
  LaneInt r;
  for (int i = 0; i < LPE; ++i) {
    float f = (float)x.v[i];
    memcpy(&r.v[i], &f, 4);
  }
  return r;
}
inline void lane_fence() {}
}  // namespace dyros
