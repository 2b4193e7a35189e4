// Forward-mode dual numbers for the tangent propagation of the EKF-RK step.
//
// The reference obtains T = J * P_sqrt with n forward-mode JVPs of the whole RK step
// (reference: src/utils.py:72-79 `jmp_aux`, called from src/filters/sqrt_ekf.py:161-163).
// Here the same forward-mode derivative is carried by a value + K tangent lanes that live
// in registers; one ODE right-hand side written once as a template serves the primal
// (T = double) and the tangent (T = Dual<K>) evaluation.
#pragma once
#include <cmath>

namespace odeu {

#ifdef ODEU_HOSTEMU
// test-only host compilation of the kernel source (tests/host_emu.cu): no device code is generated
#define ODEU_HD __host__ inline
#else
#define ODEU_HD __host__ __device__ __forceinline__
#endif

ODEU_HD double copysign_hd(double mag, double sgn) {
#ifdef __CUDA_ARCH__
  return copysign(mag, sgn);
#else
  return std::copysign(mag, sgn);
#endif
}

// 1/x and 1/sqrt(x) for NORMAL positive x without the slow-path branches of the CUDA library
// versions (which cost ~12 and ~10 instructions plus a reconvergence scope each; nine of each per
// Lorenz step): hardware seed (MUFU.RCP64H / MUFU.RSQ64H, ~20 bits) + Newton to ~1 ulp.
// x = 0 gives inf (rcp) / NaN (rsqrt after the correction): callers select around it.
ODEU_HD double rcp_pos(double x) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
#else
  return 1.0 / x;
#endif
}
ODEU_HD double rsqrt_pos(double x) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double t = x * y;
  const double e = fma(-t, y, 1.0);            // 1 - x y^2
  const double q = e * fma(0.375, e, 0.5);     // third-order correction
  return fma(y, q, y);
#else
  return 1.0 / std::sqrt(x);
#endif
}

template <int K>
struct Dual {
  double v;
  double d[K];
  ODEU_HD Dual() {}
  ODEU_HD Dual(double c) : v(c) {
#pragma unroll
    for (int k = 0; k < K; ++k) d[k] = 0.0;
  }
};

template <class T> struct is_dual { static constexpr bool value = false; };
template <int K> struct is_dual<Dual<K>> { static constexpr bool value = true; };

// primal value of either scalar kind
ODEU_HD double primal(double a) { return a; }
template <int K> ODEU_HD double primal(const Dual<K>& a) { return a.v; }

// ---- Dual (+,-) Dual
template <int K> ODEU_HD Dual<K> operator+(const Dual<K>& a, const Dual<K>& b) {
  Dual<K> r; r.v = a.v + b.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = a.d[k] + b.d[k];
  return r;
}
template <int K> ODEU_HD Dual<K> operator-(const Dual<K>& a, const Dual<K>& b) {
  Dual<K> r; r.v = a.v - b.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = a.d[k] - b.d[k];
  return r;
}
template <int K> ODEU_HD Dual<K> operator-(const Dual<K>& a) {
  Dual<K> r; r.v = -a.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = -a.d[k];
  return r;
}
// ---- Dual (+,-) double
template <int K> ODEU_HD Dual<K> operator+(const Dual<K>& a, double b) { Dual<K> r = a; r.v = a.v + b; return r; }
template <int K> ODEU_HD Dual<K> operator+(double a, const Dual<K>& b) { Dual<K> r = b; r.v = a + b.v; return r; }
template <int K> ODEU_HD Dual<K> operator-(const Dual<K>& a, double b) { Dual<K> r = a; r.v = a.v - b; return r; }
template <int K> ODEU_HD Dual<K> operator-(double a, const Dual<K>& b) {
  Dual<K> r; r.v = a - b.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = -b.d[k];
  return r;
}
// ---- products
template <int K> ODEU_HD Dual<K> operator*(const Dual<K>& a, const Dual<K>& b) {
  Dual<K> r; r.v = a.v * b.v;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = fma(a.v, b.d[k], a.d[k] * b.v);
  return r;
}
template <int K> ODEU_HD Dual<K> operator*(const Dual<K>& a, double b) {
  Dual<K> r; r.v = a.v * b;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = a.d[k] * b;
  return r;
}
template <int K> ODEU_HD Dual<K> operator*(double a, const Dual<K>& b) { return b * a; }
// ---- quotients
template <int K> ODEU_HD Dual<K> operator/(const Dual<K>& a, const Dual<K>& b) {
  Dual<K> r;
  const double ib = 1.0 / b.v;
  r.v = a.v * ib;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) * ib;
  return r;
}
template <int K> ODEU_HD Dual<K> operator/(const Dual<K>& a, double b) {
  Dual<K> r; r.v = a.v / b;
  const double ib = 1.0 / b;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = a.d[k] * ib;
  return r;
}
template <int K> ODEU_HD Dual<K> operator/(double a, const Dual<K>& b) {
  Dual<K> r;
  const double ib = 1.0 / b.v;
  r.v = a * ib;
  const double s = -r.v * ib;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = s * b.d[k];
  return r;
}

// ---- elementary functions (chain rule with the scalar derivative computed once)
ODEU_HD double d_exp(double a) { return exp(a); }
ODEU_HD double d_sin(double a) { return sin(a); }
ODEU_HD double d_cos(double a) { return cos(a); }
template <int K> ODEU_HD Dual<K> d_exp(const Dual<K>& a) {
  Dual<K> r; r.v = exp(a.v);
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = r.v * a.d[k];
  return r;
}
template <int K> ODEU_HD Dual<K> d_sin(const Dual<K>& a) {
  Dual<K> r; double s, c; sincos(a.v, &s, &c); r.v = s;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = c * a.d[k];
  return r;
}
template <int K> ODEU_HD Dual<K> d_cos(const Dual<K>& a) {
  Dual<K> r; double s, c; sincos(a.v, &s, &c); r.v = c;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = -s * a.d[k];
  return r;
}
// small integer powers as repeated products (jnp `x**2`, `x**3`, `x**4` lower to integer_pow,
// i.e. repeated multiplication; reference: src/ode/van_der_pol.py:42, src/ode/lcao.py:57,
// src/ode/hodgkin_huxley.py:46-51)
template <class T> ODEU_HD T sqr(const T& a) { return a * a; }
template <class T> ODEU_HD T cube(const T& a) { return a * a * a; }
template <class T> ODEU_HD T pow4(const T& a) { T s = a * a; return s * s; }

}  // namespace odeu
