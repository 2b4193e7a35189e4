// Negative log-likelihood AND its parameter gradient by forward-mode tangents, fused with the
// EKF time loop (BASELINE config 3: process-noise tempering for parameter estimation).
//
// Replaces `jax.value_and_grad(nll)` as driven by jaxopt.ScipyBoundedMinimize
// (scripts/run_parameter_estimation.py:599, nll :685-796).  One thread owns one
// (parameter set b, chunk of PC parameter directions): it runs the complete filter on the scalar
// type S = GDual<double, PC> (value + PC directional derivatives), so x', P', S', K' and the
// NLL derivative follow from the same statements as the value.  The RK Jacobian uses an outer
// dual over S (KC state columns per pass); its mixed part is d J / d theta.
// Every chunk recomputes the value part; chunk 0 writes the NLL.
//
// Outputs: nll [B], grad [p_opt][B] = d NLL / d theta_j for the requested flat parameter indices
// (builder order).  The host scales by (max - min) and permutes to JAX's sorted-key order
// (SURVEY Q13).
#pragma once
#include "ekf_core.cuh"
#include "gdual.cuh"
#include "dirk.cuh"

namespace odeu {

constexpr int ODEU_MAX_GRAD = 32;

template <int NX, int NP>
struct GradArgs {
  long long B, T;
  double t0, h;
  int L, noise_mode, cov_fn;
  double cov_scale;
  int ys_per_traj, has_obs;
  int p_opt;                     // number of differentiated parameters
  int idx[ODEU_MAX_GRAD];        // their flat indices (builder order)
  const double* x0;              // [n][B]
  const double* x0_tan;          // [p_opt][n][B] d x0 / d theta_j, or null (= 0)
  const double* theta;           // [NP][B] or null
  // per-trajectory diagonal Q_sqrt = diag(w) of parameter_sensitivity
  // (run_parameter_estimation.py:750-769) and its parameter tangents; replaces GQ when set
  const double* qdiag;           // [n][B] or null
  const double* qdiag_tan;       // [p_opt][n][B] or null
  double q_gamma;                // gamma_sqrt
  const double* ys; const unsigned char* flags; const long long* ymap;
  double* nll;                   // [B]
  double* grad;                  // [p_opt][B]
  double* xT;                    // [n][B] nullable (value part of the final mean)
  // full output contract of unroll() for the row kernel's NLL-only instantiation (scripts/run_filter.py:
  // 219-222): strided trajectory slots, last error estimate / pre-update y_hat, S / time, resume covariance
  long long save_interval;       // 0: none
  const double* P0b;             // [n*n][B] per-trajectory initial covariance, or null (= P0s)
  double* epsT; double* yhatT; double* ST; double* tT;
  double* out_t; double* out_x; double* out_eps; double* out_P; double* out_yhat; double* out_S;
  double P0s[NX * NX], GQ[NX * NX], H[NX * NX], R[NX * NX];
  double theta_shared[NP];
  // run-time copy of the tableau (constant-bank operands of the row kernel)
  int rt_S;
  double rt_A[8][8], rt_b[2][8], rt_c[8];
  // stage-tangent schedule of the row-parallel kernel (ekf_rows.cuh): which stage tangents the
  // propagated solution needs and where each lives (>= 0 shared-memory slot, -1 registers,
  // -2 last needed stage: accumulated straight into J)
  int rw_need[8], rw_slot[8], rw_last;
  double rw_ha[8][8], rw_hb1[8];   // h a_ij, h b1_j
  // measurement matrix with unit rows (every matrix the reference ships): observed component of
  // each row, valid when h_sel_all != 0
  int h_sel[8], h_sel_all;
};

// One RK step on scalar type S carrying KC tangent columns c0.. (identity seeds).
template <class Ode, class Tab, int KC, class S>
ODEU_HD void rk_step_generic(double t, double h, const S* x, const S* th, int c0, bool want_primal,
                             S* xn, S* eps, S (*Jcols)[KC]) {
  constexpr int n = Ode::NX;
  constexpr int St = Tab::S;
  using D = GDual<S, KC>;
  D X[n];
  for (int m = 0; m < n; ++m) {
    X[m].v = x[m];
#pragma unroll
    for (int k = 0; k < KC; ++k) X[m].d[k] = S((m == c0 + k) ? 1.0 : 0.0);
  }
  D Ks[St][n];
#pragma unroll
  for (int i = 0; i < St; ++i) {
    D Xi[n];
    for (int m = 0; m < n; ++m) {
      D s = D(0.0);
      bool any = false;
#pragma unroll
      for (int j = 0; j < i; ++j) {
        if (Tab::a(i, j) != 0.0) { s = any ? s + Ks[j][m] * Tab::a(i, j) : Ks[j][m] * Tab::a(i, j); any = true; }
      }
      Xi[m] = any ? X[m] + s * h : X[m];
    }
    Ode::rhs(t + h * Tab::c(i), Xi, th, Ks[i]);
  }
  for (int m = 0; m < n; ++m) {
    D s1 = D(0.0);
    S s0 = S(0.0);
#pragma unroll
    for (int j = 0; j < St; ++j) {
      if (Tab::b(1, j) != 0.0) s1 = s1 + Ks[j][m] * Tab::b(1, j);
      if (Tab::b(0, j) != 0.0) s0 = s0 + Ks[j][m].v * Tab::b(0, j);
    }
    const D X1 = X[m] + s1 * h;
#pragma unroll
    for (int k = 0; k < KC; ++k) Jcols[m][k] = X1.d[k];
    if (want_primal) {
      const S x0 = x[m] + s0 * h;
      xn[m] = X1.v;
      eps[m] = d_abs(x0 - X1.v);
    }
  }
}

// Generic measurement update on scalar type S (run-time L, dense constant H and R).
template <int n, class S>
ODEU_HD S correct_step_generic(int L, const double* H, const double* R, const double* y, S* x,
                               S (*P)[n]) {
  S PHt[n][n], d[n], Sm[n][n], Ls[n][n], inv[n], z[n], K[n][n];
  for (int l = 0; l < L; ++l) {
    S s = S(0.0);
    for (int j = 0; j < n; ++j) s = s + x[j] * H[l * n + j];
    d[l] = y[l] - s;
    for (int i = 0; i < n; ++i) {
      S a = S(0.0);
      for (int j = 0; j < n; ++j) a = a + P[i][j] * H[l * n + j];
      PHt[i][l] = a;
    }
  }
  for (int l = 0; l < L; ++l)
    for (int m = 0; m <= l; ++m) {
      S s = S(R[l * L + m]);
      for (int j = 0; j < n; ++j) s = s + PHt[j][m] * H[l * n + j];
      Sm[l][m] = s;
      Sm[m][l] = s;
    }
  bool all_tiny = true;
  S logdet = S(0.0), quad = S(0.0);
  for (int j = 0; j < L; ++j) {
    S s = Sm[j][j];
    for (int k = 0; k < j; ++k) s = s - Ls[j][k] * Ls[j][k];
    const S dj = d_sqrt(s);
    Ls[j][j] = dj;
    inv[j] = 1.0 / dj;
    all_tiny = all_tiny && (fabs(value_of(dj)) < 1e-16);
    logdet = logdet + d_log(d_abs(dj));
    for (int i = j + 1; i < L; ++i) {
      S v = Sm[i][j];
      for (int k = 0; k < j; ++k) v = v - Ls[i][k] * Ls[j][k];
      v = v * inv[j];
      Ls[i][j] = v;
      all_tiny = all_tiny && (fabs(value_of(v)) < 1e-16);
    }
  }
  for (int i = 0; i < L; ++i) {
    S s = d[i];
    for (int k = 0; k < i; ++k) s = s - Ls[i][k] * z[k];
    z[i] = s * inv[i];
    quad = quad + z[i] * z[i];
  }
  const S nlg = quad * 0.5 + logdet + 0.5 * (double)L * 1.8378770664093453;
  for (int i = 0; i < n; ++i) {
    S w[n];
    for (int l = 0; l < L; ++l) {
      S s = PHt[i][l];
      for (int k = 0; k < l; ++k) s = s - Ls[l][k] * w[k];
      w[l] = s * inv[l];
    }
    for (int l = L - 1; l >= 0; --l) {
      S s = w[l];
      for (int k = l + 1; k < L; ++k) s = s - Ls[k][l] * K[i][k];
      K[i][l] = all_tiny ? S(0.0) : s * inv[l];
    }
  }
  for (int i = 0; i < n; ++i) {
    S s = x[i];
    for (int l = 0; l < L; ++l) s = s + K[i][l] * d[l];
    x[i] = s;
  }
  S AP[n][n], G[n][n];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      S s = P[i][j];
      for (int l = 0; l < L; ++l) s = s - K[i][l] * PHt[j][l];
      AP[i][j] = s;
    }
  for (int i = 0; i < n; ++i)
    for (int l = 0; l < L; ++l) {
      S s = S(0.0);
      for (int j = 0; j < n; ++j) s = s + AP[i][j] * H[l * n + j];
      for (int m = 0; m < L; ++m) s = s - K[i][m] * R[m * L + l];
      G[i][l] = s;
    }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      S s = AP[i][j];
      for (int l = 0; l < L; ++l) s = s - G[i][l] * K[j][l];
      P[i][j] = s;
      P[j][i] = s;
    }
  return nlg;
}

template <class Ode, class Tab, int KC, int PC>
ODEU_HD void ekf_grad_trajectory(const GradArgs<Ode::NX, Ode::NP>& a, const long long b, const int chunk) {
  constexpr int n = Ode::NX;
  constexpr int NP = Ode::NP;
  using S = GDual<double, PC>;
  const long long B = a.B;
  const int L = a.L;
  S x[n], eps[n], P[n][n], th[NP];
  for (int i = 0; i < n; ++i) {
    x[i] = S(a.x0[i * B + b]);
#pragma unroll
    for (int q = 0; q < PC; ++q) {
      const int j = chunk * PC + q;
      if (a.x0_tan && j < a.p_opt) x[i].d[q] = a.x0_tan[((long long)j * n + i) * B + b];
    }
    eps[i] = S(0.0);
    for (int k = 0; k < n; ++k) P[i][k] = S(a.P0s[i * n + k]);
  }
  for (int k = 0; k < NP; ++k) {
    th[k] = S(a.theta ? a.theta[k * B + b] : a.theta_shared[k]);
#pragma unroll
    for (int q = 0; q < PC; ++q) {
      const int j = chunk * PC + q;
      if (j < a.p_opt && a.idx[j] == k) th[k].d[q] = 1.0;
    }
  }
  S qd[n];                       // (gamma_sqrt w_i)^2 when the per-trajectory diagonal Q is given
  for (int i = 0; i < n; ++i) {
    qd[i] = S(0.0);
    if (a.qdiag) {
      S wv = S(a.qdiag[i * B + b]);
#pragma unroll
      for (int q = 0; q < PC; ++q) {
        const int j = chunk * PC + q;
        if (a.qdiag_tan && j < a.p_opt) wv.d[q] = a.qdiag_tan[((long long)j * n + i) * B + b];
      }
      wv = wv * a.q_gamma;
      qd[i] = wv * wv;
    }
  }
  double t = a.t0;
  const double h = a.h;
  S nll = S(0.0);
  for (long long step = 0; step < a.T; ++step) {
    S xn[n], J[n][n];
    if constexpr (is_implicit<Tab>::value) {
      dirk_step_generic<Ode, Tab, S>(t, h, x, th, xn, eps, J);
    } else if constexpr (KC == n) {
      rk_step_generic<Ode, Tab, KC, S>(t, h, x, th, 0, true, xn, eps, J);
    } else {
#pragma unroll 1
      for (int c0 = 0; c0 < n; c0 += KC) {
        S Jc[n][KC];
        rk_step_generic<Ode, Tab, KC, S>(t, h, x, th, c0, c0 == 0, xn, eps, Jc);
        for (int i = 0; i < n; ++i)
#pragma unroll
          for (int k = 0; k < KC; ++k)
            if (c0 + k < n) J[i][c0 + k] = Jc[i][k];
      }
    }
    // P <- J P J^T + Q
    {
      S M[n][n];
      for (int i = 0; i < n; ++i)
        for (int k = 0; k < n; ++k) {
          S s = J[i][0] * P[0][k];
          for (int j = 1; j < n; ++j) s = s + J[i][j] * P[j][k];
          M[i][k] = s;
        }
      for (int i = 0; i < n; ++i)
        for (int k = 0; k <= i; ++k) {
          S s = M[i][0] * J[k][0];
          for (int j = 1; j < n; ++j) s = s + M[i][j] * J[k][j];
          P[i][k] = s;
          P[k][i] = s;
        }
    }
    if (a.noise_mode == NOISE_COVFN) {
      if (a.cov_fn == COV_DIAGONAL) {
        for (int i = 0; i < n; ++i) { const S e = eps[i] * a.cov_scale; P[i][i] = P[i][i] + e * e; }
      } else if (a.cov_fn == COV_OUTER) {
        for (int i = 0; i < n; ++i)
          for (int k = 0; k < n; ++k) P[i][k] = P[i][k] + (eps[i] * a.cov_scale) * (eps[k] * a.cov_scale);
      } else {
        for (int i = 0; i < n; ++i) P[i][i] = P[i][i] + a.cov_scale * a.cov_scale;
      }
    } else if (a.noise_mode == NOISE_EPS_PLUS_Q) {
      for (int i = 0; i < n; ++i) {
        if (a.qdiag) P[i][i] = P[i][i] + qd[i];
        else for (int k = 0; k < n; ++k) P[i][k] = P[i][k] + a.GQ[i * n + k];
        P[i][i] = P[i][i] + eps[i] * eps[i];
      }
    } else if (a.noise_mode == NOISE_Q_ONLY) {
      for (int i = 0; i < n; ++i) {
        if (a.qdiag) P[i][i] = P[i][i] + qd[i];
        else for (int k = 0; k < n; ++k) P[i][k] = P[i][k] + a.GQ[i * n + k];
      }
    }
    for (int i = 0; i < n; ++i) x[i] = xn[i];
    t = t + h;
    if (a.has_obs && a.flags[step]) {
      const long long oi = a.ymap[step];
      double y[n];
      for (int l = 0; l < L; ++l) y[l] = a.ys_per_traj ? a.ys[(oi * L + l) * B + b] : a.ys[oi * L + l];
      nll = nll + correct_step_generic<n, S>(L, a.H, a.R, y, x, P);
    }
  }
  if (chunk == 0) {
    if (a.nll) a.nll[b] = nll.v;
    if (a.xT)
      for (int i = 0; i < n; ++i) a.xT[i * B + b] = x[i].v;
  }
#pragma unroll
  for (int q = 0; q < PC; ++q) {
    const int j = chunk * PC + q;
    if (j < a.p_opt && a.grad) a.grad[(long long)j * B + b] = nll.d[q];
  }
}

// grid: x = trajectory blocks, y = direction chunks
template <class Ode, class Tab, int KC, int PC, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
ekf_grad_kernel(const __grid_constant__ GradArgs<Ode::NX, Ode::NP> a) {
  const long long b = (long long)blockIdx.x * BLOCK + threadIdx.x;
  if (b >= a.B) return;
  ekf_grad_trajectory<Ode, Tab, KC, PC>(a, b, (int)blockIdx.y);
}

}  // namespace odeu
