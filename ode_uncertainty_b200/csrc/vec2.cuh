// Two-wide scalar for the row-parallel kernel: one thread carries the SAME equation of TWO
// trajectories, every arithmetic statement expands to two independent instructions in the same
// basic block.  The row kernel is bound by the latency of its per-warp dependency chains (a lone CTA
// takes as long per step as two co-resident ones take together), so the second
// chain fills issue slots that are otherwise idle.  Also works as the component type of GDual
// (one-lane dual in V inside Ode::row).
#pragma once
#include "gdual.cuh"

namespace odeu {

struct V2d {
  double a, b;
  ODEU_HD V2d() {}
  ODEU_HD V2d(double c) : a(c), b(c) {}
  ODEU_HD V2d(double x, double y) : a(x), b(y) {}
};

ODEU_HD V2d operator+(const V2d& x, const V2d& y) { return V2d(x.a + y.a, x.b + y.b); }
ODEU_HD V2d operator-(const V2d& x, const V2d& y) { return V2d(x.a - y.a, x.b - y.b); }
ODEU_HD V2d operator*(const V2d& x, const V2d& y) { return V2d(x.a * y.a, x.b * y.b); }
ODEU_HD V2d operator/(const V2d& x, const V2d& y) { return V2d(x.a / y.a, x.b / y.b); }
ODEU_HD V2d operator-(const V2d& x) { return V2d(-x.a, -x.b); }
ODEU_HD V2d operator+(const V2d& x, double y) { return V2d(x.a + y, x.b + y); }
ODEU_HD V2d operator+(double x, const V2d& y) { return V2d(x + y.a, x + y.b); }
ODEU_HD V2d operator-(const V2d& x, double y) { return V2d(x.a - y, x.b - y); }
ODEU_HD V2d operator-(double x, const V2d& y) { return V2d(x - y.a, x - y.b); }
ODEU_HD V2d operator*(const V2d& x, double y) { return V2d(x.a * y, x.b * y); }
ODEU_HD V2d operator*(double x, const V2d& y) { return V2d(x * y.a, x * y.b); }
ODEU_HD V2d operator/(const V2d& x, double y) { return V2d(x.a / y, x.b / y); }
ODEU_HD V2d operator/(double x, const V2d& y) { return V2d(x / y.a, x / y.b); }
ODEU_HD V2d d_exp(const V2d& x) { return V2d(exp(x.a), exp(x.b)); }
ODEU_HD V2d d_sqrt(const V2d& x) { return V2d(sqrt(x.a), sqrt(x.b)); }
ODEU_HD V2d d_rsqrt(const V2d& x) { return V2d(rsqrt_pos(x.a), rsqrt_pos(x.b)); }
ODEU_HD V2d d_fma(const V2d& x, const V2d& y, const V2d& z) { return V2d(fma(x.a, y.a, z.a), fma(x.b, y.b, z.b)); }
ODEU_HD V2d d_log(const V2d& x) { return V2d(log(x.a), log(x.b)); }
ODEU_HD V2d d_abs(const V2d& x) { return V2d(fabs(x.a), fabs(x.b)); }

// ---- trajectories per scalar and per-lane access (row kernel bookkeeping)
template <class S> struct lanes_of { static constexpr int value = 1; };
template <> struct lanes_of<V2d> { static constexpr int value = 2; };

ODEU_HD double lane_get(double s, int) { return s; }
ODEU_HD double lane_get(const V2d& s, int u) { return u == 0 ? s.a : s.b; }
template <class T, int K> ODEU_HD double lane_get(const GDual<T, K>& s, int u) { return lane_get(s.v, u); }
ODEU_HD void lane_set(double& s, int, double v) { s = v; }
ODEU_HD void lane_set(V2d& s, int u, double v) { if (u == 0) s.a = v; else s.b = v; }
template <class T, int K> ODEU_HD void lane_set(GDual<T, K>& s, int u, double v) { lane_set(s.v, u, v); }

// per-lane predicate "numerically zero" and the masked select used by the zero-gain guard
struct Mask2 { bool m[2]; };
ODEU_HD Mask2 mask_true() { Mask2 r; r.m[0] = r.m[1] = true; return r; }
ODEU_HD void mask_and_tiny(Mask2& k, double s) { k.m[0] = k.m[0] && (fabs(s) < 1e-16); k.m[1] = k.m[0]; }
ODEU_HD void mask_and_tiny(Mask2& k, const V2d& s) {
  k.m[0] = k.m[0] && (fabs(s.a) < 1e-16);
  k.m[1] = k.m[1] && (fabs(s.b) < 1e-16);
}
template <class T, int K> ODEU_HD void mask_and_tiny(Mask2& k, const GDual<T, K>& s) { mask_and_tiny(k, s.v); }
// per-lane "exactly zero" (an exactly singular innovation: pivot 0, column 0 instead of 0 * inf)
ODEU_HD Mask2 mask_zero(double s) { Mask2 r; r.m[0] = r.m[1] = (s == 0.0); return r; }
ODEU_HD Mask2 mask_zero(const V2d& s) { Mask2 r; r.m[0] = (s.a == 0.0); r.m[1] = (s.b == 0.0); return r; }
template <class T, int K> ODEU_HD Mask2 mask_zero(const GDual<T, K>& s) { return mask_zero(s.v); }
ODEU_HD double zero_where(const Mask2& k, double s) { return k.m[0] ? 0.0 : s; }
ODEU_HD V2d zero_where(const Mask2& k, const V2d& s) { return V2d(k.m[0] ? 0.0 : s.a, k.m[1] ? 0.0 : s.b); }
template <class T, int K> ODEU_HD GDual<T, K> zero_where(const Mask2& k, const GDual<T, K>& s) {
  return k.m[0] ? GDual<T, K>(0.0) : s;       // gradient scalars carry one trajectory
}

}  // namespace odeu
