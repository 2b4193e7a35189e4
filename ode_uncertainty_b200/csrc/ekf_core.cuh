// Per-trajectory EKF-over-embedded-RK step: device building blocks shared by all kernels.
//
// What one step computes (reference call stack: scripts/run_filter.py:204-217 `unroll`,
// scripts/run_parameter_estimation.py:782-793 `nll`):
//   predict  src/filters/sqrt_ekf.py:92-197   RK step + J = d x_next / d x by forward mode,
//                                             P <- J P J^T + Q(eps, gamma, Q_sqrt)
//   correct  src/filters/sqrt_ekf.py:337-376  S = H P H^T + R, K = P H^T S^-1 (Cholesky solve),
//                                             x <- x + K (y - H x), Joseph update of P
//   nll      src/utils.py:109-128             0.5 |Ls^-1 (y - yhat)|^2 + L/2 log 2pi + sum log Ls_ii
//
// Covariance representation: the reference is a square-root filter whose factor is defined only
// up to column signs; this implementation carries the full symmetric P, which is the
// sign-free quantity the reference's own tests compare (tests/test_utils.py:31).  The zero-gain
// guard `all(S_sqrt < 1e-16)` (sqrt_ekf.py:351-353) is implemented with its intended meaning
// `all(|chol(S)| < 1e-16)`; the oracle reports every step where the LAPACK-sign
// version would differ.
#pragma once
#include <cstring>
#include "dual.cuh"
#include "tableaux.cuh"
#include "odes.cuh"

namespace odeu {

// noise branch table of SURVEY section 3.2 (src/filters/sqrt_ekf.py:96-136,172-180)
enum NoiseMode {
  NOISE_COVFN = 0,     // disable_cov_update=false, no Q:   P += covfn(eps)
  NOISE_EPS_PLUS_Q = 1,  // disable_cov_update=false, Q any:  P += diag(eps^2) + gamma Q Q^T
  NOISE_Q_ONLY = 2,    // disable_cov_update=true,  Q any:  P += gamma Q Q^T
  NOISE_NONE = 3       // disable_cov_update=true,  no Q:   P unchanged
};
enum CovFn { COV_DIAGONAL = 0, COV_OUTER = 1, COV_STATIC_DIAGONAL = 2 };

// ---------------------------------------------------------------------------------------------
// One explicit embedded RK step carrying KC tangent columns (identity seeds c0..c0+KC-1).
// Restates src/solvers/rksolver.py:113-155 (+ compute_node :160-194):
//   x_i = x + h * (ks @ A[i]),  k_i = f(t + h c_i, x_i)
//   x_next[r] = x + h * (ks @ b[r]),  propagate r=1,  eps = |x_next[0] - x_next[1]|
// The sums skip the structural zeros of A and b; x_next[0] - x_next[1] is evaluated as the
// difference of the two full solutions, like the reference, so eps keeps its cancellation.
//
// Tangent lanes carry no cancellation, so they accumulate straight into the identity seed with
// the pre-scaled coefficients ha[i][j] = h a_ij, hb1[j] = h b_1j (kernel arguments, i.e.
// constant-bank operands of the DFMAs): one DFMA per non-zero coefficient, no separate h-scale.
// `seed` (optional, [n][KC]): tangent seeds instead of identity columns - the factor form seeds
// with the columns of P_sqrt like jmp_aux (src/utils.py:72-79), so Jcols returns T = J P_sqrt.
struct ScaledTableau {
  double ha[8][8];
  double hb1[8];
  // unscaled coefficients of the primal sums: as kernel arguments they are constant-bank operands
  // of the DFMAs; as compile-time immediates each cost two UMOVs per use (50 per Lorenz step)
  double a[8][8];
  double b0[8], b1[8];
};

template <class Ode, class Tab, int KC, class PT>
ODEU_HD void rk_step_tangent(double t, double h, const ScaledTableau& st, const double* x,
                             const PT* th, int c0, bool want_primal, double* xn, double* eps,
                             double (*Jcols)[KC], const double (*seed)[KC] = nullptr) {
  constexpr int n = Ode::NX;
  constexpr int S = Tab::S;
  using D = Dual<KC>;
  D X[n];
#pragma unroll
  for (int m = 0; m < n; ++m) {
    X[m].v = x[m];
#pragma unroll
    for (int k = 0; k < KC; ++k) X[m].d[k] = seed ? seed[m][k] : ((m == c0 + k) ? 1.0 : 0.0);
  }
  D Ks[S][n];
#pragma unroll
  for (int i = 0; i < S; ++i) {
    D Xi[n];
    if (i == 0) {
#pragma unroll
      for (int m = 0; m < n; ++m) Xi[m] = X[m];
    } else {
#pragma unroll
      for (int m = 0; m < n; ++m) {
        double sv = 0.0;
        bool first = true;
        Xi[m] = X[m];
#pragma unroll
        for (int j = 0; j < i; ++j) {
          if (Tab::a(i, j) != 0.0) {
            sv = first ? Ks[j][m].v * st.a[i][j] : fma(st.a[i][j], Ks[j][m].v, sv);
            first = false;
#pragma unroll
            for (int k = 0; k < KC; ++k) Xi[m].d[k] = fma(st.ha[i][j], Ks[j][m].d[k], Xi[m].d[k]);
          }
        }
        if (!first) Xi[m].v = fma(h, sv, X[m].v);
      }
    }
    Ode::rhs(t + h * Tab::c(i), Xi, th, Ks[i]);
  }
#pragma unroll
  for (int m = 0; m < n; ++m) {
    double s1v = 0.0, s0 = 0.0;
    bool f1 = true, f0 = true;
#pragma unroll
    for (int k = 0; k < KC; ++k) Jcols[m][k] = X[m].d[k];
#pragma unroll
    for (int j = 0; j < S; ++j) {
      if (Tab::b(1, j) != 0.0) {
        s1v = f1 ? Ks[j][m].v * st.b1[j] : fma(st.b1[j], Ks[j][m].v, s1v);
        f1 = false;
#pragma unroll
        for (int k = 0; k < KC; ++k) Jcols[m][k] = fma(st.hb1[j], Ks[j][m].d[k], Jcols[m][k]);
      }
      if (Tab::b(0, j) != 0.0) {
        if (f0) { s0 = Ks[j][m].v * st.b0[j]; f0 = false; }
        else s0 = fma(st.b0[j], Ks[j][m].v, s0);
      }
    }
    if (want_primal) {
      const double x1 = fma(h, s1v, x[m]);
      const double x0 = fma(h, s0, x[m]);
      xn[m] = x1;
      eps[m] = fabs(x0 - x1);
    }
  }
}

// Plain RK step without tangents (particle ensemble, src/filters/particle_filter.py:87).
template <class Ode, class Tab, class PT>
ODEU_HD void rk_step_plain(double t, double h, const double* x, const PT* th,
                                              double* xn, double* eps) {
  constexpr int n = Ode::NX;
  constexpr int S = Tab::S;
  double Ks[S][n];
#pragma unroll
  for (int i = 0; i < S; ++i) {
    double Xi[n];
#pragma unroll
    for (int m = 0; m < n; ++m) {
      double s = 0.0;
      bool first = true;
#pragma unroll
      for (int j = 0; j < i; ++j) {
        if (Tab::a(i, j) != 0.0) {
          if (first) { s = Ks[j][m] * Tab::a(i, j); first = false; }
          else s = fma(Tab::a(i, j), Ks[j][m], s);
        }
      }
      Xi[m] = first ? x[m] : fma(h, s, x[m]);
    }
    Ode::rhs(t + h * Tab::c(i), Xi, th, Ks[i]);
  }
#pragma unroll
  for (int m = 0; m < n; ++m) {
    double s1 = 0.0, s0 = 0.0;
    bool f1 = true, f0 = true;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      if (Tab::b(1, j) != 0.0) {
        if (f1) { s1 = Ks[j][m] * Tab::b(1, j); f1 = false; }
        else s1 = fma(Tab::b(1, j), Ks[j][m], s1);
      }
      if (Tab::b(0, j) != 0.0) {
        if (f0) { s0 = Ks[j][m] * Tab::b(0, j); f0 = false; }
        else s0 = fma(Tab::b(0, j), Ks[j][m], s0);
      }
    }
    const double x1 = fma(h, s1, x[m]);
    const double x0 = fma(h, s0, x[m]);
    xn[m] = x1;
    eps[m] = fabs(x0 - x1);
  }
}

// Same step, same operation order (so the same bits), with the tableau coefficients read from the kernel arguments:
// as constant-bank operands of the DFMAs they cost nothing, as compile-time immediates two UMOVs per use (ncu: 40 of
// the 370 instructions of a Lorenz particle-step).
template <class Ode, class Tab, class PT>
ODEU_HD void rk_step_plain_st(double t, double h, const ScaledTableau& st, const double* x, const PT* th,
                              double* xn, double* eps) {
  constexpr int n = Ode::NX;
  constexpr int S = Tab::S;
  double Ks[S][n];
#pragma unroll
  for (int i = 0; i < S; ++i) {
    double Xi[n];
#pragma unroll
    for (int m = 0; m < n; ++m) {
      double s = 0.0;
      bool first = true;
#pragma unroll
      for (int j = 0; j < i; ++j) {
        if (Tab::a(i, j) != 0.0) {
          if (first) { s = Ks[j][m] * st.a[i][j]; first = false; }
          else s = fma(st.a[i][j], Ks[j][m], s);
        }
      }
      Xi[m] = first ? x[m] : fma(h, s, x[m]);
    }
    Ode::rhs(t + h * Tab::c(i), Xi, th, Ks[i]);
  }
#pragma unroll
  for (int m = 0; m < n; ++m) {
    double s1 = 0.0, s0 = 0.0;
    bool f1 = true, f0 = true;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      if (Tab::b(1, j) != 0.0) {
        if (f1) { s1 = Ks[j][m] * st.b1[j]; f1 = false; }
        else s1 = fma(st.b1[j], Ks[j][m], s1);
      }
      if (Tab::b(0, j) != 0.0) {
        if (f0) { s0 = Ks[j][m] * st.b0[j]; f0 = false; }
        else s0 = fma(st.b0[j], Ks[j][m], s0);
      }
    }
    const double x1 = fma(h, s1, x[m]);
    const double x0 = fma(h, s0, x[m]);
    xn[m] = x1;
    eps[m] = fabs(x0 - x1);
  }
}

// ---------------------------------------------------------------------------------------------
// P <- J P J^T (symmetric result; both triangles written so later code can index freely).
template <int n>
ODEU_HD void propagate_cov(const double (*J)[n], double (*P)[n]) {
  constexpr int U = (n <= 4) ? n : 1;
  double M[n][n];
#pragma unroll U
  for (int i = 0; i < n; ++i)
#pragma unroll U
    for (int k = 0; k < n; ++k) {
      double s = J[i][0] * P[0][k];
#pragma unroll U
      for (int j = 1; j < n; ++j) s = fma(J[i][j], P[j][k], s);
      M[i][k] = s;
    }
#pragma unroll U
  for (int i = 0; i < n; ++i)
#pragma unroll U
    for (int k = 0; k <= i; ++k) {
      double s = M[i][0] * J[k][0];
#pragma unroll U
      for (int j = 1; j < n; ++j) s = fma(M[i][j], J[k][j], s);
      P[i][k] = s;
      P[k][i] = s;
    }
}

// Process-noise branch table (src/filters/sqrt_ekf.py:96-136; covariance-update plugins
// src/covariance_update_functions/{diagonal.py:43-58, outer.py:44-62, static_diagonal.py:31-48}).
template <int n>
ODEU_HD void add_process_noise(int noise_mode, int cov_fn, double scale,
                                                  const double* eps, const double* GQ,
                                                  double (*P)[n]) {
  constexpr int U = (n <= 4) ? n : 1;
  if (noise_mode == NOISE_COVFN) {
    if (cov_fn == COV_DIAGONAL) {
#pragma unroll U
      for (int i = 0; i < n; ++i) { const double e = scale * eps[i]; P[i][i] = fma(e, e, P[i][i]); }
    } else if (cov_fn == COV_OUTER) {
      // Q_sqrt = outer(se,se)/|se| in the reference (outer.py:57-60): 0/0 = NaN when eps == 0
      double ss = 0.0;
#pragma unroll U
      for (int i = 0; i < n; ++i) { const double e = scale * eps[i]; ss = fma(e, e, ss); }
      const double poison = (ss == 0.0) ? (ss / ss) : 0.0;  // NaN
#pragma unroll U
      for (int i = 0; i < n; ++i)
#pragma unroll U
        for (int k = 0; k < n; ++k)
          P[i][k] = fma(scale * eps[i], scale * eps[k], P[i][k]) + poison;
    } else {  // static diagonal: P += c^2 I
#pragma unroll U
      for (int i = 0; i < n; ++i) P[i][i] = fma(scale, scale, P[i][i]);
    }
  } else if (noise_mode == NOISE_EPS_PLUS_Q) {
    // cov_update_fn and its scale are ignored on this branch (sqrt_ekf.py:97-101)
#pragma unroll U
    for (int i = 0; i < n; ++i)
#pragma unroll U
      for (int k = 0; k < n; ++k) P[i][k] += GQ[i * n + k];
#pragma unroll U
    for (int i = 0; i < n; ++i) P[i][i] = fma(eps[i], eps[i], P[i][i]);
  } else if (noise_mode == NOISE_Q_ONLY) {
#pragma unroll U
    for (int i = 0; i < n; ++i)
#pragma unroll U
      for (int k = 0; k < n; ++k) P[i][k] += GQ[i * n + k];
  }
}

// Where a measurement update publishes the PRE-update y_hat and S that the reference keeps in
// its state dict (sqrt_ekf.py:370-372).  They are written straight to the caller's output
// buffers (trajectory slot and/or final-state arrays; batch-minor, stride B) instead of being
// carried in registers for the whole run.
struct ObsSink {
  double* y1; double* S1;   // upcoming trajectory slot (nullable)
  double* y2; double* S2;   // final-state outputs (nullable)
  long long stride;
  bool any;                 // some pointer is set (warp-uniform): lets the caller branch around the
                            // whole publication instead of issuing 2 (L + L^2) predicated-off stores
  ODEU_HD void put_y(int l, double v) const {
    if (y1) y1[l * stride] = v;
    if (y2) y2[l * stride] = v;
  }
  ODEU_HD void put_S(int idx, double v) const {
    if (S1) S1[idx * stride] = v;
    if (S2) S2[idx * stride] = v;
  }
};

// ---------------------------------------------------------------------------------------------
// Measurement update + log-likelihood term.  L (<= n) is a run-time value; all loops are
// statically bounded by n and guarded so small systems stay in registers.
// Writes the PRE-update y_hat and S like the reference state (sqrt_ekf.py:370-372).
// Returns the negative log Gaussian term of this observation (src/utils.py:109-128).
template <int n>
ODEU_HD double correct_step(int L, const double* H, const double* R,
                                               const double* y, double* x, double (*P)[n],
                                               const ObsSink& sink) {
  constexpr int U = (n <= 4) ? n : 1;
  double PHt[n][n];  // [n][L]
  double d[n];
  double Smat[n][n];
  // y_hat = H x, PHt = P H^T
#pragma unroll U
  for (int l = 0; l < n; ++l) {
    if (l < L) {
      double s = 0.0;
#pragma unroll U
      for (int j = 0; j < n; ++j) s = fma(H[l * n + j], x[j], s);
      sink.put_y(l, s);
      d[l] = y[l] - s;
#pragma unroll U
      for (int i = 0; i < n; ++i) {
        double a = 0.0;
#pragma unroll U
        for (int j = 0; j < n; ++j) a = fma(P[i][j], H[l * n + j], a);
        PHt[i][l] = a;
      }
    }
  }
  // S = H PHt + R  (symmetric)
#pragma unroll U
  for (int l = 0; l < n; ++l)
#pragma unroll U
    for (int m = 0; m <= l; ++m) {
      if (l < L) {
        double s = R[l * L + m];
#pragma unroll U
        for (int j = 0; j < n; ++j) s = fma(H[l * n + j], PHt[j][m], s);
        Smat[l][m] = s;
        Smat[m][l] = s;
        sink.put_S(l * L + m, s);
        if (m != l) sink.put_S(m * L + l, s);
      }
    }
  // Cholesky S = Ls Ls^T; the reciprocal diagonal is kept so every later triangular solve
  // multiplies instead of dividing (FP64 division costs ~10x a DFMA on the FP64 pipe)
  double Ls[n][n], inv[n];
  bool all_tiny = true;
#pragma unroll U
  for (int j = 0; j < n; ++j) {
    if (j < L) {
      double s = Smat[j][j];
#pragma unroll U
      for (int k = 0; k < j; ++k) s = fma(-Ls[j][k], Ls[j][k], s);
      const double dj = sqrt(s);
      Ls[j][j] = dj;
      all_tiny = all_tiny && (fabs(dj) < 1e-16);
      inv[j] = 1.0 / dj;
#pragma unroll U
      for (int i = j + 1; i < n; ++i) {
        if (i < L) {
          double v = Smat[i][j];
#pragma unroll U
          for (int k = 0; k < j; ++k) v = fma(-Ls[i][k], Ls[j][k], v);
          v = (dj == 0.0) ? 0.0 : v * inv[j];    // exactly singular S: a zero column, not 0 * inf
          Ls[i][j] = v;
          all_tiny = all_tiny && (fabs(v) < 1e-16);
        }
      }
    }
  }
  // z = Ls^-1 d ; nlg
  double z[n];
  double quad = 0.0, logdet = 0.0;
#pragma unroll U
  for (int i = 0; i < n; ++i) {
    if (i < L) {
      double s = d[i];
#pragma unroll U
      for (int k = 0; k < i; ++k) s = fma(-Ls[i][k], z[k], s);
      z[i] = s * inv[i];
      quad = fma(z[i], z[i], quad);
      logdet += log(fabs(Ls[i][i]));
    }
  }
  const double nlg = 0.5 * quad + 0.5 * (double)L * 1.8378770664093453 + logdet;

  // K = PHt S^-1 : solve row by row, K[i][:] = (Ls^-T Ls^-1 PHt[i][:])
  double K[n][n];  // [n][L]
  if (all_tiny) {
#pragma unroll U
    for (int i = 0; i < n; ++i)
#pragma unroll U
      for (int l = 0; l < n; ++l) K[i][l] = 0.0;
  } else {
#pragma unroll U
    for (int i = 0; i < n; ++i) {
      double w[n];
#pragma unroll U
      for (int l = 0; l < n; ++l) {
        if (l < L) {
          double s = PHt[i][l];
#pragma unroll U
          for (int k = 0; k < l; ++k) s = fma(-Ls[l][k], w[k], s);
          w[l] = s * inv[l];
        }
      }
#pragma unroll U
      for (int l = n - 1; l >= 0; --l) {
        if (l < L) {
          double s = w[l];
#pragma unroll U
          for (int k = l + 1; k < n; ++k)
            if (k < L) s = fma(-Ls[k][l], K[i][k], s);
          K[i][l] = s * inv[l];
        }
      }
    }
  }
  // x <- x + K d
#pragma unroll U
  for (int i = 0; i < n; ++i) {
    double s = x[i];
#pragma unroll U
    for (int l = 0; l < n; ++l)
      if (l < L) s = fma(K[i][l], d[l], s);
    x[i] = s;
  }
  // Joseph: A = I - K H;  P <- A P A^T + K R K^T
  //   AP = P - K PHt^T ;  APHt = AP H^T ;  P+ = AP - APHt K^T + (K R) K^T
  double AP[n][n];
#pragma unroll U
  for (int i = 0; i < n; ++i)
#pragma unroll U
    for (int j = 0; j < n; ++j) {
      double s = P[i][j];
#pragma unroll U
      for (int l = 0; l < n; ++l)
        if (l < L) s = fma(-K[i][l], PHt[j][l], s);
      AP[i][j] = s;
    }
  double G[n][n];  // [n][L] : APHt - K R   (so that P+ = AP - G K^T)
#pragma unroll U
  for (int i = 0; i < n; ++i)
#pragma unroll U
    for (int l = 0; l < n; ++l) {
      if (l < L) {
        double s = 0.0;
#pragma unroll U
        for (int j = 0; j < n; ++j) s = fma(AP[i][j], H[l * n + j], s);
#pragma unroll U
        for (int m = 0; m < n; ++m)
          if (m < L) s = fma(-K[i][m], R[m * L + l], s);
        G[i][l] = s;
      }
    }
#pragma unroll U
  for (int i = 0; i < n; ++i)
#pragma unroll U
    for (int j = 0; j <= i; ++j) {
      double s = AP[i][j];
#pragma unroll U
      for (int l = 0; l < n; ++l)
        if (l < L) s = fma(-G[i][l], K[j][l], s);
      P[i][j] = s;
      P[j][i] = s;
    }
  return nlg;
}

// ---------------------------------------------------------------------------------------------
// Measurement update specialised at compile time for H = [I_L 0] (the first L state components
// are observed directly) - the shape of every measurement_matrix the reference ships for small
// systems ('[[1, 0]]', identity, ...; configs/*/*.yaml).  Same mathematics as correct_step:
// H P = first L rows of P, so the H products disappear; Cholesky pivots use rsqrt, all solves
// multiply by the stored reciprocals, and sum(log Ls_ii) = 0.5 log(prod pivots) needs one log.
// Running product of Cholesky pivots kept as mantissa x 2^exponent: sum_t sum_i log Ls_ii =
// 0.5 log(prod_t prod_i pivot), so the FP64 `log` (about 35 instructions) is taken once per time
// segment instead of once per step; per step it costs one DMUL and a few integer operations.
struct LogProd {
  double m;
  int e;
  ODEU_HD static long long to_bits(double v) {
#ifdef __CUDA_ARCH__
    return __double_as_longlong(v);
#else
    long long b; memcpy(&b, &v, sizeof(b)); return b;
#endif
  }
  ODEU_HD static double from_bits(long long b) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double(b);
#else
    double v; memcpy(&v, &b, sizeof(v)); return v;
#endif
  }
  ODEU_HD void reset() { m = 1.0; e = 0; }
  ODEU_HD void mul(double piv) {            // piv: normal, positive (checked by the caller)
    const long long bits = to_bits(m * piv);
    e += (int)((bits >> 52) & 0x7ff) - 1023;
    m = from_bits((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);   // mantissa in [1, 2)
  }
  ODEU_HD double flush() {                  // 0.5 log(product), then start over
    const double r = 0.5 * (log(m) + (double)e * 0.6931471805599453);
    reset();
    return r;
  }
};

template <int n, int L>
ODEU_HD double correct_step_lead(const double* R, const double* y, double* x, double (*P)[n],
                                 const ObsSink& sink, LogProd* lp = nullptr) {
  double d[L], Ls[L][L], inv[L], Smat[L][L];
  // publication of the pre-update y_hat / S: rare on throughput runs (only the last observation step
  // of a run writes the final-state buffers), so it sits behind ONE uniform branch - as predicated
  // stores it cost ~100 of the ~1,000 instructions of a Lorenz step (ncu source view, round 1g)
  if (sink.any) {
#pragma unroll
    for (int l = 0; l < L; ++l) {
      sink.put_y(l, x[l]);
#pragma unroll
      for (int m = 0; m <= l; ++m) {
        const double s = P[l][m] + R[l * L + m];
        sink.put_S(l * L + m, s);
        if (m != l) sink.put_S(m * L + l, s);
      }
    }
  }
#pragma unroll
  for (int l = 0; l < L; ++l) {
    d[l] = y[l] - x[l];
#pragma unroll
    for (int m = 0; m <= l; ++m) Smat[l][m] = P[l][m] + R[l * L + m];
  }
  bool all_tiny = true;
  double piv = 1.0;
#pragma unroll
  for (int j = 0; j < L; ++j) {
    double s = Smat[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) s = fma(-Ls[j][k], Ls[j][k], s);
    piv *= s;
    const double r = rsqrt_pos(s);
    const double dj = (s == 0.0) ? 0.0 : s * r;   // sqrt(s) up to an ulp; NaN for s < 0 like sqrt; an exactly
                                         // singular S (P = 0, R = 0) must trip the guard, not poison it (0 * inf)
    inv[j] = r;
    Ls[j][j] = dj;
    all_tiny = all_tiny && (fabs(dj) < 1e-16);
#pragma unroll
    for (int i = j + 1; i < L; ++i) {
      double v = Smat[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v = fma(-Ls[i][k], Ls[j][k], v);
      v = (s == 0.0) ? 0.0 : v * r;        // exactly singular S: a zero column, not 0 * inf
      Ls[i][j] = v;
      all_tiny = all_tiny && (fabs(v) < 1e-16);
    }
  }
  double z[L], quad = 0.0;
#pragma unroll
  for (int i = 0; i < L; ++i) {
    double s = d[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s = fma(-Ls[i][k], z[k], s);
    z[i] = s * inv[i];
    quad = fma(z[i], z[i], quad);
  }
  // sum_i log|Ls_ii| = 0.5 log(prod_i pivot_i); fall back to the sum when the product leaves
  // the normal range (or a pivot is non-positive / NaN)
  double logdet;
  if (piv > 1e-290 && piv < 1e290) {
    if (lp) { lp->mul(piv); logdet = 0.0; }     // log deferred to LogProd::flush()
    else logdet = 0.5 * log(piv);
  } else {
    logdet = 0.0;
#pragma unroll
    for (int i = 0; i < L; ++i) logdet += log(fabs(Ls[i][i]));
  }
  const double nlg = 0.5 * quad + 0.5 * (double)L * 1.8378770664093453 + logdet;

  double K[n][L];
  if (all_tiny) {
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int l = 0; l < L; ++l) K[i][l] = 0.0;
  } else {
#pragma unroll
    for (int i = 0; i < n; ++i) {
      double w[L];
#pragma unroll
      for (int l = 0; l < L; ++l) {
        double s = P[i][l];
#pragma unroll
        for (int k = 0; k < l; ++k) s = fma(-Ls[l][k], w[k], s);
        w[l] = s * inv[l];
      }
#pragma unroll
      for (int l = L - 1; l >= 0; --l) {
        double s = w[l];
#pragma unroll
        for (int k = l + 1; k < L; ++k) s = fma(-Ls[k][l], K[i][k], s);
        K[i][l] = s * inv[l];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < n; ++i) {
    double s = x[i];
#pragma unroll
    for (int l = 0; l < L; ++l) s = fma(K[i][l], d[l], s);
    x[i] = s;
  }
  // Joseph with H = [I_L 0]:  AP = P - K P[:L,:];  G = AP[:, :L] - K R;  P+ = AP - G K^T
  double AP[n][n];
#pragma unroll
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int j = 0; j < n; ++j) {
      double s = P[i][j];
#pragma unroll
      for (int l = 0; l < L; ++l) s = fma(-K[i][l], P[l][j], s);
      AP[i][j] = s;
    }
  double G[n][L];
#pragma unroll
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int l = 0; l < L; ++l) {
      double s = AP[i][l];
#pragma unroll
      for (int m = 0; m < L; ++m) s = fma(-K[i][m], R[m * L + l], s);
      G[i][l] = s;
    }
#pragma unroll
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double s = AP[i][j];
#pragma unroll
      for (int l = 0; l < L; ++l) s = fma(-G[i][l], K[j][l], s);
      P[i][j] = s;
      P[j][i] = s;
    }
  return nlg;
}

}  // namespace odeu
