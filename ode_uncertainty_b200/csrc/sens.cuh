// Parameter-sensitivity weights for the process noise (scripts/run_parameter_estimation.py:750-769):
//   jac  = jacfwd(theta -> solver(ode, theta, initial_state).x)(theta)     one RK step from (t0, x0)
//   w_i  = sum over the OPTIMISED scalar parameters k of |d x1_i / d theta_k|
//   w    = sqrt(n) * w / |w|_2 ,   Q_sqrt = diag(w)
// The reference evaluates this INSIDE the differentiated loss, so the optimiser's gradient contains
// d w / d theta_j: second derivatives of the step (and, with initial_state_parametrized, the
// dependence of x0 on theta).  One thread = (parameter set b, direction j): inner dual S1 carries
// direction j (theta_j seeded 1, x0 carries d x0 / d theta_j), the outer dual seeds theta_k for the
// PARTIAL derivative at fixed x0 (the closure `solver_jac_params_wrapper` captures x0 as a constant).
#pragma once
#include "dirk.cuh"
#include "ekf_grad.cuh"

namespace odeu {

template <int NP>
struct SensArgs {
  long long B;
  double t0, h;
  int p_opt;
  int idx[ODEU_MAX_GRAD];
  const double* x0;        // [n][B]
  const double* x0_tan;    // [p_opt][n][B] or null
  const double* theta;     // [NP][B] or null
  double* w;               // [n][B]
  double* w_tan;           // [p_opt][n][B] or null
  double theta_shared[NP];
};

// Plain RK step (propagating row b[1], rksolver.py:146-151) on an arbitrary scalar type.
template <class Ode, class Tab, class D>
ODEU_HD void rk_step_scalar(double t, double h, const D* x, const D* th, D* xn) {
  if constexpr (is_implicit<Tab>::value) {       // implicit solver plugins (dirk.cuh)
    D eps_[Ode::NX], J_[Ode::NX][Ode::NX];
    dirk_step_generic<Ode, Tab, D>(t, h, x, th, xn, eps_, J_);
    return;
  } else {
  constexpr int n = Ode::NX;
  constexpr int St = Tab::S;
  D Ks[St][n];
#pragma unroll 1
  for (int i = 0; i < St; ++i) {
    D Xi[n];
    for (int m = 0; m < n; ++m) {
      D s = D(0.0);
      for (int j = 0; j < i; ++j)
        if (Tab::a(i, j) != 0.0) s = s + Ks[j][m] * Tab::a(i, j);
      Xi[m] = x[m] + s * h;
    }
    Ode::rhs(t + h * Tab::c(i), Xi, th, Ks[i]);
  }
  for (int m = 0; m < n; ++m) {
    D s = D(0.0);
    for (int j = 0; j < St; ++j)
      if (Tab::b(1, j) != 0.0) s = s + Ks[j][m] * Tab::b(1, j);
    xn[m] = x[m] + s * h;
  }
  }
}

template <class Ode, class Tab>
ODEU_HD void param_sens_unit(const SensArgs<Ode::NP>& a, const long long b, const int j) {
  constexpr int n = Ode::NX;
  constexpr int NP = Ode::NP;
  using S1 = GDual<double, 1>;
  using D = GDual<S1, 1>;
  const long long B = a.B;
  S1 w[n];
  for (int i = 0; i < n; ++i) w[i] = S1(0.0);
#pragma unroll 1
  for (int k = 0; k < a.p_opt; ++k) {
    D x[n], th[NP], xn[n];
    for (int i = 0; i < n; ++i) {
      S1 v = S1(a.x0[i * B + b]);
      if (a.x0_tan) v.d[0] = a.x0_tan[((long long)j * n + i) * B + b];
      x[i] = D(v);
    }
    for (int q = 0; q < NP; ++q) {
      S1 v = S1(a.theta ? a.theta[q * B + b] : a.theta_shared[q]);
      if (a.idx[j] == q) v.d[0] = 1.0;
      th[q] = D(v);
      if (a.idx[k] == q) th[q].d[0] = S1(1.0);
    }
    rk_step_scalar<Ode, Tab, D>(a.t0, a.h, x, th, xn);
    for (int i = 0; i < n; ++i) w[i] = w[i] + d_abs(xn[i].d[0]);
  }
  S1 ss = S1(0.0);
  for (int i = 0; i < n; ++i) ss = ss + w[i] * w[i];
  const S1 nrm = d_sqrt(ss);
  const double rn = sqrt((double)n);
  for (int i = 0; i < n; ++i) {
    const S1 v = (w[i] * rn) / nrm;
    if (j == 0) a.w[i * B + b] = v.v;
    if (a.w_tan) a.w_tan[((long long)j * n + i) * B + b] = v.d[0];
  }
}

template <class Ode, class Tab>
__global__ void __launch_bounds__(64) param_sens_kernel(const SensArgs<Ode::NP> a) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < a.B) param_sens_unit<Ode, Tab>(a, b, (int)blockIdx.y);
}

}  // namespace odeu
