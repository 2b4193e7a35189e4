// Generic (nestable) forward-mode dual numbers for the parameter-gradient kernels.
//
// The reference differentiates the NLL of the whole scan in reverse mode
// (jaxopt.ScipyBoundedMinimize -> jax.value_and_grad(nll), scripts/run_parameter_estimation.py:599).
// Here the same derivative is carried forward, per parameter direction, through the whole EKF
// step: the inner scalar S = GDual<double, PC> holds a value and its derivative along PC
// parameter directions; the RK Jacobian J = d x_next / d x is obtained, exactly like in the
// plain filter, from an OUTER dual over S (GDual<S, KC>), whose mixed part is the second-order
// term d J / d theta that forward mode needs.
#pragma once
#include <cmath>
#include <type_traits>

#include "dual.cuh"

namespace odeu {

template <class S, int K>
struct GDual {
  S v;
  S d[K];
  ODEU_HD GDual() {}
  ODEU_HD GDual(double c) : v(c) {
#pragma unroll
    for (int k = 0; k < K; ++k) d[k] = S(0.0);
  }
  template <class U = S, class = typename std::enable_if<!std::is_same<U, double>::value>::type>
  ODEU_HD GDual(const S& c) : v(c) {
#pragma unroll
    for (int k = 0; k < K; ++k) d[k] = S(0.0);
  }
};

template <class T> struct is_gdual : std::false_type {};
template <class S, int K> struct is_gdual<GDual<S, K>> : std::true_type {};

// innermost value
ODEU_HD double value_of(double a) { return a; }
template <class S, int K> ODEU_HD double value_of(const GDual<S, K>& a) { return value_of(a.v); }

#define ODEU_NOT_DOUBLE(S) typename std::enable_if<!std::is_same<S, double>::value, int>::type = 0

// ---- GDual (op) GDual
template <class S, int K> ODEU_HD GDual<S, K> operator+(const GDual<S, K>& a, const GDual<S, K>& b) {
  GDual<S, K> r; r.v = a.v + b.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = a.d[k] + b.d[k];
  return r;
}
template <class S, int K> ODEU_HD GDual<S, K> operator-(const GDual<S, K>& a, const GDual<S, K>& b) {
  GDual<S, K> r; r.v = a.v - b.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = a.d[k] - b.d[k];
  return r;
}
template <class S, int K> ODEU_HD GDual<S, K> operator-(const GDual<S, K>& a) {
  GDual<S, K> r; r.v = -a.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = -a.d[k];
  return r;
}
template <class S, int K> ODEU_HD GDual<S, K> operator*(const GDual<S, K>& a, const GDual<S, K>& b) {
  GDual<S, K> r; r.v = a.v * b.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = a.v * b.d[k] + a.d[k] * b.v;
  return r;
}
template <class S, int K> ODEU_HD GDual<S, K> operator/(const GDual<S, K>& a, const GDual<S, K>& b) {
  GDual<S, K> r;
  const S ib = 1.0 / b.v;
  r.v = a.v * ib;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) * ib;
  return r;
}
// ---- GDual (op) S  and  S (op) GDual   (S = the component type, possibly double)
template <class S, int K> ODEU_HD GDual<S, K> operator+(const GDual<S, K>& a, const S& b) { GDual<S, K> r = a; r.v = a.v + b; return r; }
template <class S, int K> ODEU_HD GDual<S, K> operator+(const S& a, const GDual<S, K>& b) { GDual<S, K> r = b; r.v = a + b.v; return r; }
template <class S, int K> ODEU_HD GDual<S, K> operator-(const GDual<S, K>& a, const S& b) { GDual<S, K> r = a; r.v = a.v - b; return r; }
template <class S, int K> ODEU_HD GDual<S, K> operator-(const S& a, const GDual<S, K>& b) {
  GDual<S, K> r; r.v = a - b.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = -b.d[k];
  return r;
}
template <class S, int K> ODEU_HD GDual<S, K> operator*(const GDual<S, K>& a, const S& b) {
  GDual<S, K> r; r.v = a.v * b;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = a.d[k] * b;
  return r;
}
template <class S, int K> ODEU_HD GDual<S, K> operator*(const S& a, const GDual<S, K>& b) { return b * a; }
template <class S, int K> ODEU_HD GDual<S, K> operator/(const GDual<S, K>& a, const S& b) {
  GDual<S, K> r;
  const S ib = 1.0 / b;
  r.v = a.v * ib;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = a.d[k] * ib;
  return r;
}
template <class S, int K> ODEU_HD GDual<S, K> operator/(const S& a, const GDual<S, K>& b) {
  GDual<S, K> r;
  const S ib = 1.0 / b.v;
  r.v = a * ib;
  const S s = -(r.v * ib);
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = s * b.d[k];
  return r;
}
// ---- GDual (op) double when the component type is itself a dual
template <class S, int K, ODEU_NOT_DOUBLE(S)> ODEU_HD GDual<S, K> operator+(const GDual<S, K>& a, double b) { GDual<S, K> r = a; r.v = a.v + b; return r; }
template <class S, int K, ODEU_NOT_DOUBLE(S)> ODEU_HD GDual<S, K> operator+(double a, const GDual<S, K>& b) { GDual<S, K> r = b; r.v = a + b.v; return r; }
template <class S, int K, ODEU_NOT_DOUBLE(S)> ODEU_HD GDual<S, K> operator-(const GDual<S, K>& a, double b) { GDual<S, K> r = a; r.v = a.v - b; return r; }
template <class S, int K, ODEU_NOT_DOUBLE(S)> ODEU_HD GDual<S, K> operator-(double a, const GDual<S, K>& b) {
  GDual<S, K> r; r.v = a - b.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = -b.d[k];
  return r;
}
template <class S, int K, ODEU_NOT_DOUBLE(S)> ODEU_HD GDual<S, K> operator*(const GDual<S, K>& a, double b) {
  GDual<S, K> r; r.v = a.v * b;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = a.d[k] * b;
  return r;
}
template <class S, int K, ODEU_NOT_DOUBLE(S)> ODEU_HD GDual<S, K> operator*(double a, const GDual<S, K>& b) { return b * a; }
template <class S, int K, ODEU_NOT_DOUBLE(S)> ODEU_HD GDual<S, K> operator/(const GDual<S, K>& a, double b) { return a * (1.0 / b); }
template <class S, int K, ODEU_NOT_DOUBLE(S)> ODEU_HD GDual<S, K> operator/(double a, const GDual<S, K>& b) {
  GDual<S, K> r;
  const S ib = 1.0 / b.v;
  r.v = a * ib;
  const S s = -(r.v * ib);
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = s * b.d[k];
  return r;
}

// ---- fused multiply-add a b + c: value 1 FMA, each direction 2 FMAs (the operator form a * b + c costs
// MUL + FMA + ADD per direction, because the sum of the two cross terms cannot be re-associated into c)
ODEU_HD double d_fma(double a, double b, double c) { return fma(a, b, c); }
template <class S, int K> ODEU_HD GDual<S, K> d_fma(const GDual<S, K>& a, const GDual<S, K>& b, const GDual<S, K>& c) {
  GDual<S, K> r; r.v = d_fma(a.v, b.v, c.v);
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = d_fma(a.v, b.d[k], d_fma(a.d[k], b.v, c.d[k]));
  return r;
}

// ---- elementary functions (recursive in the component type)
template <class S, int K> ODEU_HD GDual<S, K> d_exp(const GDual<S, K>& a) {
  GDual<S, K> r; r.v = d_exp(a.v);
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = r.v * a.d[k];
  return r;
}
template <class S, int K> ODEU_HD GDual<S, K> d_sin(const GDual<S, K>& a) {
  GDual<S, K> r; r.v = d_sin(a.v);
  const S c = d_cos(a.v);
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = c * a.d[k];
  return r;
}
template <class S, int K> ODEU_HD GDual<S, K> d_cos(const GDual<S, K>& a) {
  GDual<S, K> r; r.v = d_cos(a.v);
  const S s = -d_sin(a.v);
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = s * a.d[k];
  return r;
}
ODEU_HD double d_sqrt(double a) { return sqrt(a); }
// 1 / sqrt(a) for a normal positive a (no slow path, see rsqrt_pos); derivative -a' / (2 a sqrt(a))
ODEU_HD double d_rsqrt(double a) { return rsqrt_pos(a); }

ODEU_HD double d_log(double a) { return log(a); }
ODEU_HD double d_abs(double a) { return fabs(a); }
template <class S, int K> ODEU_HD GDual<S, K> d_sqrt(const GDual<S, K>& a) {
  GDual<S, K> r; r.v = d_sqrt(a.v);
  const S h = 0.5 / r.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = h * a.d[k];
  return r;
}
template <class S, int K> ODEU_HD GDual<S, K> d_rsqrt(const GDual<S, K>& a) {
  GDual<S, K> r; r.v = d_rsqrt(a.v);
  const S h = -0.5 * r.v * r.v * r.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = h * a.d[k];
  return r;
}
template <class S, int K> ODEU_HD GDual<S, K> d_log(const GDual<S, K>& a) {
  GDual<S, K> r; r.v = d_log(a.v);
  const S i = 1.0 / a.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = i * a.d[k];
  return r;
}
// |a| with derivative sign(a) a' (0 at a == 0, like jax.lax.abs's JVP)
template <class S, int K> ODEU_HD GDual<S, K> d_abs(const GDual<S, K>& a) {
  const double av = value_of(a);
  const double sg = (av > 0.0) ? 1.0 : ((av < 0.0) ? -1.0 : 0.0);
  GDual<S, K> r; r.v = d_abs(a.v);
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = a.d[k] * sg;
  return r;
}

}  // namespace odeu
