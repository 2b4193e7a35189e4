// Sign-faithful square-root form of the covariance steps (plan guard mode "reference").
//
// The reference is a square-root filter: every covariance operation is `sqrt_L_sum_qr(a, b)` =
// R^T of the economic QR of [a^T; b^T] (src/utils.py:233-274), evaluated by LAPACK's Householder
// QR, whose diagonal sign follows dlarfg (beta = -sign(alpha) * norm; H = I when the sub-column is
// exactly zero).  Those signs are observable: the zero-gain guard `all(S_sqrt < 1e-16)`
// (src/filters/sqrt_ekf.py:350-353) is true for a healthy but all-negative factor and then drops
// the observation.  The full-covariance kernels cannot see the signs, so this file
// carries the factor itself through the three QRs of a step
//   predict  sqrt_ekf.py:96-136   P_sqrt <- qr([T^T; Q-blocks]).R^T, T = J P_sqrt by forward mode
//   correct  sqrt_ekf.py:349      S_sqrt <- qr([(H P_sqrt)^T; R_sqrt^T]).R^T
//   Joseph   sqrt_ekf.py:359-360  P_sqrt <- qr([((I - K H) P_sqrt)^T; (K R_sqrt)^T]).R^T
// with the same reflector sequence as dgeqr2, so the sign pattern (and therefore the guard) is the
// reference's.  Magnitudes agree with LAPACK to rounding (the reflector is applied unnormalised:
// H = I + w w^T / (beta u), w = [alpha - beta; x], one rsqrt and one reciprocal per column).
#pragma once
#include "ekf_core.cuh"

namespace odeu {

// Structure of a stacked matrix, known at compile time: first(r) = first column in which row r can
// be non-zero on entry (a row joins the reflector of column j once first(r) <= j; NC = never);
// single(r) = the row holds exactly one non-zero, at column first(r) (diagonal blocks).
struct DenseRows {
  ODEU_HD static constexpr int first(int) { return 0; }
  ODEU_HD static constexpr bool single(int) { return false; }
};

// In-place Householder triangularisation of A [MR][NC] (rows >= NC are eliminated), R only.
// On return the upper triangle of the first NC rows holds R with dgeqr2/dlarfg signs; rinv[j] =
// 1 / R_jj.  ncols: columns actually present (run-time, <= NC).
// Branch-free per column: a sub-column that is exactly zero (dlarfg: H = I, beta = alpha) selects
// g = 0, so the trailing update adds zeros, instead of branching around it.
template <int MR, int NC, class St>
ODEU_HD void householder_R(double (&A)[MR][NC], double* rinv, int ncols = NC) {
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    if (j < ncols) {
      double xn2 = 0.0;
#pragma unroll
      for (int i = j + 1; i < MR; ++i)
        if (St::first(i) <= j) xn2 = fma(A[i][j], A[i][j], xn2);
      const double alpha = A[j][j];
      const bool ident = !(xn2 != 0.0);                    // (a NaN sub-column takes the reflector path)
      const double s2 = fma(alpha, alpha, xn2);
      const double r = rsqrt_pos(s2);
      const double nrm = s2 * r;
      // Fortran SIGN: the sign BIT of alpha decides (alpha = +0 reflects to -norm)
      const double beta = ident ? alpha : -copysign_hd(nrm, alpha);
      const double u = alpha - beta;
      // H = I + w w^T / (beta u), w = [u; x];  beta u = -(|alpha| nrm + s2): no cancellation
      const double g = ident ? 0.0 : -rcp_pos(fma(fabs(alpha), nrm, s2));
#pragma unroll
      for (int k = j + 1; k < NC; ++k) {
        if (k < ncols) {
          double wa = u * A[j][k];
#pragma unroll
          for (int i = j + 1; i < MR; ++i)
            if (St::first(i) <= j && !(St::single(i) && St::first(i) == j)) wa = fma(A[i][j], A[i][k], wa);
          wa *= g;
          A[j][k] = fma(wa, u, A[j][k]);
#pragma unroll
          for (int i = j + 1; i < MR; ++i) {
            if (St::first(i) <= j) {
              if (St::single(i) && St::first(i) == j) A[i][k] = wa * A[i][j];
              else A[i][k] = fma(wa, A[i][j], A[i][k]);
            }
          }
        }
      }
      A[j][j] = beta;
      rinv[j] = ident ? copysign_hd(r, alpha) : -copysign_hd(r, alpha);   // 1 / R_jj (r = 1/|alpha| when ident)
    }
  }
}

// P = Ps Ps^T (what the ABI exchanges)
template <int n>
ODEU_HD void factor_to_cov(const double (*Ps)[n], double (*P)[n]) {
  constexpr int U = (n <= 4) ? n : 1;
#pragma unroll U
  for (int i = 0; i < n; ++i)
#pragma unroll U
    for (int j = 0; j <= i; ++j) {
      double s = 0.0;
#pragma unroll U
      for (int k = 0; k < n; ++k) s = fma(Ps[i][k], Ps[j][k], s);
      P[i][j] = s;
      P[j][i] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// Predict, structured: noise branch NOISE_COVFN with a diagonal second block (Diagonal or
// StaticDiagonal plugin): qr([T^T; diag(e)]).  Result lower-triangular.
template <int n>
struct PredictDiagRows {          // rows 0..n-1: T^T (dense); row n+d: e_d at column d
  ODEU_HD static constexpr int first(int r) { return r < n ? 0 : r - n; }
  ODEU_HD static constexpr bool single(int r) { return r >= n; }
};

template <int n>
ODEU_HD void predict_factor_diag(const double (*T)[n], const double* e, double (*Ps)[n]) {
  double A[2 * n][n], rinv[n];
#pragma unroll
  for (int r = 0; r < n; ++r)
#pragma unroll
    for (int c = 0; c < n; ++c) {
      A[r][c] = T[c][r];
      A[n + r][c] = (r == c) ? e[r] : 0.0;
    }
  householder_R<2 * n, n, PredictDiagRows<n>>(A, rinv);
#pragma unroll
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int j = 0; j < n; ++j) Ps[i][j] = (j <= i) ? A[j][i] : 0.0;
}

// Predict, generic: any noise branch (sqrt_ekf.py:96-136); block ORDER as in the reference.
//   NOISE_COVFN      sqrt_L_sum_qr(T, covfn_sqrt(eps))            diagonal.py:43-58, outer.py:44-62,
//                                                                 static_diagonal.py:31-48
//   NOISE_EPS_PLUS_Q sqrt_L_sum_qr_3(g Q_sqrt, diag(eps), T)
//   NOISE_Q_ONLY     sqrt_L_sum_qr(T, g Q_sqrt)
//   NOISE_NONE       T (not triangular)
template <int n>
ODEU_HD void predict_factor_generic(int noise_mode, int cov_fn, double scale, const double* eps,
                                    const double* gQs /* gamma_sqrt * Q_sqrt, [n][n] */,
                                    const double (*T)[n], double (*Ps)[n]) {
  constexpr int U = (n <= 4) ? n : 1;
  if (noise_mode == NOISE_NONE) {
#pragma unroll U
    for (int i = 0; i < n; ++i)
#pragma unroll U
      for (int j = 0; j < n; ++j) Ps[i][j] = T[i][j];
    return;
  }
  double A[3 * n][n], rinv[n];
#pragma unroll U
  for (int r = 0; r < 3 * n; ++r)
#pragma unroll U
    for (int c = 0; c < n; ++c) A[r][c] = 0.0;
  if (noise_mode == NOISE_EPS_PLUS_Q) {
#pragma unroll U
    for (int r = 0; r < n; ++r)
#pragma unroll U
      for (int c = 0; c < n; ++c) {
        A[r][c] = gQs[c * n + r];
        A[n + r][c] = (r == c) ? eps[r] : 0.0;
        A[2 * n + r][c] = T[c][r];
      }
  } else {
#pragma unroll U
    for (int r = 0; r < n; ++r)
#pragma unroll U
      for (int c = 0; c < n; ++c) A[r][c] = T[c][r];
    if (noise_mode == NOISE_Q_ONLY) {
#pragma unroll U
      for (int r = 0; r < n; ++r)
#pragma unroll U
        for (int c = 0; c < n; ++c) A[n + r][c] = gQs[c * n + r];
    } else if (cov_fn == COV_DIAGONAL) {
#pragma unroll U
      for (int r = 0; r < n; ++r) A[n + r][r] = scale * eps[r];
    } else if (cov_fn == COV_OUTER) {      // outer(se, se) / |se|: 0/0 = NaN when eps == 0
      double ss = 0.0;
#pragma unroll U
      for (int i = 0; i < n; ++i) { const double se = scale * eps[i]; ss = fma(se, se, ss); }
      const double nrm = sqrt(ss);
#pragma unroll U
      for (int r = 0; r < n; ++r)
#pragma unroll U
        for (int c = 0; c < n; ++c) A[n + r][c] = (scale * eps[c]) * (scale * eps[r]) / nrm;
    } else {
#pragma unroll U
      for (int r = 0; r < n; ++r) A[n + r][r] = scale;
    }
  }
  householder_R<3 * n, n, DenseRows>(A, rinv);
#pragma unroll U
  for (int i = 0; i < n; ++i)
#pragma unroll U
    for (int j = 0; j < n; ++j) Ps[i][j] = (j <= i) ? A[j][i] : 0.0;
}

// What the guard did on this update (per-trajectory counters of the run)
struct GuardCount {
  int fired;      // K = 0 was taken
  int mismatch;   // verbatim `all(S_sqrt < 1e-16)` and intended `all(|S_sqrt| < 1e-16)` disagree
};

// ---------------------------------------------------------------------------------------------
// Measurement update in factor form, H = [I_L 0], Ps lower-triangular and Rs DIAGONAL on entry
// (every script of the reference builds R_sqrt = const_diag(L, sqrt(obs_noise_var))).
//   S_sqrt: column j of [Ps[:L,:]^T; Rs] has non-zeros in rows 0..j and row n+j only, so the
//   reflector of column j touches row j and the rows n..n+j of the R_sqrt block (fill-in of the
//   earlier reflectors plus r_j): alpha_j = Ps[j][j].
template <int n, int L>
struct LeadSRows {
  ODEU_HD static constexpr int first(int r) { return r < L ? r : (r < n ? L : r - n); }   // L = never
  ODEU_HD static constexpr bool single(int r) { return r >= n; }
};

template <int n, int L>
ODEU_HD double correct_factor_lead(const double* R, const double* Rs, const double* y, double* x,
                                   double (*Ps)[n], const ObsSink& sink, LogProd* lp, bool verbatim,
                                   GuardCount& gc) {
  double d[L], PHt[n][L];
#pragma unroll
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int l = 0; l < L; ++l) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < n; ++c)
        if (c <= i && c <= l) s = fma(Ps[i][c], Ps[l][c], s);
      PHt[i][l] = s;
    }
  if (sink.any) {
#pragma unroll
    for (int l = 0; l < L; ++l) {
      sink.put_y(l, x[l]);
#pragma unroll
      for (int m = 0; m <= l; ++m) {
        const double s = PHt[l][m] + R[l * L + m];
        sink.put_S(l * L + m, s);
        if (m != l) sink.put_S(m * L + l, s);
      }
    }
  }
#pragma unroll
  for (int l = 0; l < L; ++l) d[l] = y[l] - x[l];
  double M[n + L][L], inv[L];
#pragma unroll
  for (int r = 0; r < n; ++r)
#pragma unroll
    for (int c = 0; c < L; ++c) M[r][c] = (r <= c) ? Ps[c][r] : 0.0;
#pragma unroll
  for (int k = 0; k < L; ++k)
#pragma unroll
    for (int c = 0; c < L; ++c) M[n + k][c] = (c == k) ? Rs[k * L + k] : 0.0;
  householder_R<n + L, L, LeadSRows<n, L>>(M, inv);
  // S_sqrt[i][j] = M[j][i]; guards (sqrt_ekf.py:350-353: the zeros above the diagonal pass `< 1e-16`)
  bool g_ref = true, g_int = true;
  double piv = 1.0;
#pragma unroll
  for (int j = 0; j < L; ++j) {
    piv *= M[j][j] * M[j][j];
#pragma unroll
    for (int i = j; i < L; ++i) {
      g_ref = g_ref && (M[j][i] < 1e-16);
      g_int = g_int && (fabs(M[j][i]) < 1e-16);
    }
  }
  const bool zeroK = verbatim ? g_ref : g_int;
  gc.fired += zeroK ? 1 : 0;
  gc.mismatch += (g_ref != g_int) ? 1 : 0;
  double z[L], quad = 0.0;
#pragma unroll
  for (int i = 0; i < L; ++i) {
    double s = d[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s = fma(-M[k][i], z[k], s);
    z[i] = s * inv[i];
    quad = fma(z[i], z[i], quad);
  }
  double logdet;
  if (piv > 1e-290 && piv < 1e290) {
    if (lp) { lp->mul(piv); logdet = 0.0; }
    else logdet = 0.5 * log(piv);
  } else {
    logdet = 0.0;
#pragma unroll
    for (int i = 0; i < L; ++i) logdet += log(fabs(M[i][i]));
  }
  const double nlg = 0.5 * quad + 0.5 * (double)L * 1.8378770664093453 + logdet;

  double K[n][L];
  if (zeroK) {
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
        double s = PHt[i][l];
#pragma unroll
        for (int k = 0; k < l; ++k) s = fma(-M[k][l], w[k], s);
        w[l] = s * inv[l];
      }
#pragma unroll
      for (int l = L - 1; l >= 0; --l) {
        double s = w[l];
#pragma unroll
        for (int k = l + 1; k < L; ++k) s = fma(-M[l][k], K[i][k], s);
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
  // Joseph: qr([((I - K H) Ps)^T; (K Rs)^T]); (I - K H) Ps = Ps - K Ps[:L, :]
  double A2[n + L][n], rinv[n];
#pragma unroll
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int c = 0; c < n; ++c) {
      double s = (c <= i) ? Ps[i][c] : 0.0;
#pragma unroll
      for (int l = 0; l < L; ++l)
        if (c <= l) s = fma(-K[i][l], Ps[l][c], s);
      A2[c][i] = s;
    }
#pragma unroll
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int m = 0; m < L; ++m) A2[n + m][i] = K[i][m] * Rs[m * L + m];
  householder_R<n + L, n, DenseRows>(A2, rinv);
#pragma unroll
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int j = 0; j < n; ++j) Ps[i][j] = (j <= i) ? A2[j][i] : 0.0;
  return nlg;
}

// Measurement update in factor form, generic: run-time L <= n, dense H, dense Ps / Rs.
template <int n>
ODEU_HD double correct_factor_generic(int L, const double* H, const double* Rs, const double* y,
                                      double* x, double (*Ps)[n], const ObsSink& sink, bool verbatim,
                                      GuardCount& gc) {
  constexpr int U = (n <= 4) ? n : 1;
  double HP[n][n], d[n];       // HP [L][n]
#pragma unroll U
  for (int l = 0; l < n; ++l) {
    if (l < L) {
      double s = 0.0;
#pragma unroll U
      for (int j = 0; j < n; ++j) s = fma(H[l * n + j], x[j], s);
      sink.put_y(l, s);
      d[l] = y[l] - s;
#pragma unroll U
      for (int c = 0; c < n; ++c) {
        double a = 0.0;
#pragma unroll U
        for (int j = 0; j < n; ++j) a = fma(H[l * n + j], Ps[j][c], a);
        HP[l][c] = a;
      }
    }
  }
  double M[2 * n][n], inv[n];
#pragma unroll U
  for (int r = 0; r < n; ++r)
#pragma unroll U
    for (int c = 0; c < n; ++c) {
      M[r][c] = (c < L) ? HP[c][r] : 0.0;
      M[n + r][c] = (c < L && r < L) ? Rs[c * L + r] : 0.0;
    }
  householder_R<2 * n, n, DenseRows>(M, inv, L);
  bool g_ref = true, g_int = true;
  double logdet = 0.0;
#pragma unroll U
  for (int j = 0; j < n; ++j)
#pragma unroll U
    for (int i = 0; i < n; ++i)
      if (j < L && i < L && i >= j) {
        g_ref = g_ref && (M[j][i] < 1e-16);
        g_int = g_int && (fabs(M[j][i]) < 1e-16);
      }
  if (sink.any) {     // S = S_sqrt S_sqrt^T
#pragma unroll U
    for (int l = 0; l < n; ++l)
#pragma unroll U
      for (int m = 0; m < n; ++m)
        if (l < L && m <= l) {
          double s = 0.0;
#pragma unroll U
          for (int k = 0; k < n; ++k)
            if (k <= m) s = fma(M[k][l], M[k][m], s);
          sink.put_S(l * L + m, s);
          if (m != l) sink.put_S(m * L + l, s);
        }
  }
  const bool zeroK = verbatim ? g_ref : g_int;
  gc.fired += zeroK ? 1 : 0;
  gc.mismatch += (g_ref != g_int) ? 1 : 0;
  double z[n], quad = 0.0;
#pragma unroll U
  for (int i = 0; i < n; ++i) {
    if (i < L) {
      double s = d[i];
#pragma unroll U
      for (int k = 0; k < i; ++k) s = fma(-M[k][i], z[k], s);
      z[i] = s * inv[i];
      quad = fma(z[i], z[i], quad);
      logdet += log(fabs(M[i][i]));
    }
  }
  const double nlg = 0.5 * quad + 0.5 * (double)L * 1.8378770664093453 + logdet;
  // K = (cho_solve(S_sqrt, H) Ps Ps^T)^T = (Ps HP^T) S^-1
  double K[n][n];   // [n][L]
  if (zeroK) {
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
          double s = 0.0;
#pragma unroll U
          for (int c = 0; c < n; ++c) s = fma(Ps[i][c], HP[l][c], s);
#pragma unroll U
          for (int k = 0; k < l; ++k) s = fma(-M[k][l], w[k], s);
          w[l] = s * inv[l];
        }
      }
#pragma unroll U
      for (int l = n - 1; l >= 0; --l) {
        if (l < L) {
          double s = w[l];
#pragma unroll U
          for (int k = l + 1; k < n; ++k)
            if (k < L) s = fma(-M[l][k], K[i][k], s);
          K[i][l] = s * inv[l];
        }
      }
    }
  }
#pragma unroll U
  for (int i = 0; i < n; ++i) {
    double s = x[i];
#pragma unroll U
    for (int l = 0; l < n; ++l)
      if (l < L) s = fma(K[i][l], d[l], s);
    x[i] = s;
  }
  double A2[2 * n][n], rinv[n];
#pragma unroll U
  for (int i = 0; i < n; ++i)
#pragma unroll U
    for (int c = 0; c < n; ++c) {
      double s = Ps[i][c];
#pragma unroll U
      for (int l = 0; l < n; ++l)
        if (l < L) s = fma(-K[i][l], HP[l][c], s);
      A2[c][i] = s;
      double kr = 0.0;                 // (K Rs)[i][c], zero for c >= L
#pragma unroll U
      for (int l = 0; l < n; ++l)
        if (l < L && c < L) kr = fma(K[i][l], Rs[l * L + c], kr);
      A2[n + c][i] = kr;
    }
  householder_R<2 * n, n, DenseRows>(A2, rinv);
#pragma unroll U
  for (int i = 0; i < n; ++i)
#pragma unroll U
    for (int j = 0; j < n; ++j) Ps[i][j] = (j <= i) ? A2[j][i] : 0.0;
  return nlg;
}

}  // namespace odeu
