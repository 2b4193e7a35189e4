// Implicit (stiff) solver plugins: diagonally implicit Runge-Kutta steps with a Newton solve per
// stage - SURVEY 8(f) N3.
//
// The reference reaches these through `DiffraxSolverBuilder` (src/solvers/diffrax_solver.py:16-140):
// diffrax's `Kvaerno3` (or `ImplicitEuler`, the builder's default) with
// `Newton(rtol=1e-8, atol=1e-8)`, `root_find_max_steps=500`, one step of size h per call, and
// `eps = zeros` (:128) - every shipped Hodgkin-Huxley configuration selects it
// (configs/params/hodgkinhuxley*.yaml:10-14).  diffrax is a third-party dependency that is absent
// from /root/reference and from this image, so this file restates the PUBLISHED method
// (A. Kvaerno, "Singly diagonally implicit Runge-Kutta methods with an explicit first stage",
// BIT 44 (2004): ESDIRK 3(2) with 4 stages, stiffly accurate, gamma = 0.43586652150...; diffrax
// 0.6.0 uses the same tableau) and PARITY IS UNPINNED: the oracle for it is this repository's own
// restatement in oracle/ref_torch.py (`dirk_step`), not a reference output.
//
// Per implicit stage i:  z = x + h sum_{j<i} a_ij k_j + h a_ii f(t_i, z),  k_i = f(t_i, z),
// solved by full Newton (Jacobian I - h a_ii Df(z) re-evaluated every iteration, dense LU with
// partial pivoting on n <= 16) to |dz_m| <= 1e-12 (1 + |z_m|) - tighter than the reference's 1e-8 so
// that the result does not depend on the iteration count - plus ONE further iteration, after which
// the derivative lanes of the scalar type S are exact too (at a root the Newton map's derivative is
// the implicit-function derivative).  The step Jacobian J = d x_next / d x follows from the implicit
// function theorem, stage by stage:  dZ_i = (I - h a_ii Df_i)^-1 (I + h sum_j a_ij dK_j),
// dK_i = Df_i dZ_i.  Stiffly accurate tableaux: x_next = z of the last stage.
//
// S is `double` (filter / solver trajectories) or GDual<double, PC> (NLL gradient): the same
// statements carry the parameter tangents, including d J / d theta through the nested dual in the
// Jacobian evaluation.
#pragma once
#include <limits>
#include "gdual.cuh"
#include "tableaux.cuh"

namespace odeu {

struct TabKvaerno3 {
  static constexpr bool IMPLICIT = true;
  static constexpr int S = 4;
  static constexpr double g = 0.43586652150845899941601945;
  __host__ __device__ static constexpr double a(int i, int j) {
    constexpr double A[4][4] = {
        {0.0, 0.0, 0.0, 0.0},
        {g, g, 0.0, 0.0},
        {(-4.0 * g * g + 6.0 * g - 1.0) / (4.0 * g), (-2.0 * g + 1.0) / (4.0 * g), g, 0.0},
        {(6.0 * g - 1.0) / (12.0 * g), -1.0 / ((24.0 * g - 12.0) * g), (-6.0 * g * g + 6.0 * g - 1.0) / (6.0 * g - 3.0), g}};
    return A[i][j];
  }
  __host__ __device__ static constexpr double c(int i) {
    constexpr double C[4] = {0.0, 2.0 * g, 1.0, 1.0};
    return C[i];
  }
  // explicit-solver interface (never used for stepping; lets shared host code compile)
  __host__ __device__ static constexpr double b(int r, int j) { return a(3, j); }
};

struct TabImplicitEuler {        // diffrax ImplicitEuler: x_next = x + h f(t + h, x_next)
  static constexpr bool IMPLICIT = true;
  static constexpr int S = 1;
  __host__ __device__ static constexpr double a(int, int) { return 1.0; }
  __host__ __device__ static constexpr double c(int) { return 1.0; }
  __host__ __device__ static constexpr double b(int, int) { return 1.0; }
};

template <class Tab, class = void> struct is_implicit : std::false_type {};
template <class Tab> struct is_implicit<Tab, std::void_t<decltype(Tab::IMPLICIT)>> : std::true_type {};

// Solve A X = B in place (A [n][n], B [n][m]) by LU with partial pivoting on the VALUE part.
template <int n, int m, class S>
ODEU_HD void lu_solve(S (*A)[n], S (*Bm)[m]) {
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = fabs(value_of(A[k][k]));
    for (int i = k + 1; i < n; ++i) {
      const double v = fabs(value_of(A[i][k]));
      if (v > best) { best = v; p = i; }
    }
    if (p != k) {
      for (int j = 0; j < n; ++j) { const S tmp = A[k][j]; A[k][j] = A[p][j]; A[p][j] = tmp; }
      for (int j = 0; j < m; ++j) { const S tmp = Bm[k][j]; Bm[k][j] = Bm[p][j]; Bm[p][j] = tmp; }
    }
    const S piv = 1.0 / A[k][k];
    for (int i = k + 1; i < n; ++i) {
      const S f = A[i][k] * piv;
      for (int j = k + 1; j < n; ++j) A[i][j] = A[i][j] - f * A[k][j];
      for (int j = 0; j < m; ++j) Bm[i][j] = Bm[i][j] - f * Bm[k][j];
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    const S piv = 1.0 / A[k][k];
    for (int j = 0; j < m; ++j) {
      S s = Bm[k][j];
      for (int i = k + 1; i < n; ++i) s = s - A[k][i] * Bm[i][j];
      Bm[k][j] = s * piv;
    }
  }
}

// f(t, z) and Df(t, z) on scalar type S: one right-hand-side evaluation per state column with a
// one-lane dual over S (keeps the live state small; the Hodgkin-Huxley right-hand side is large).
template <class Ode, class S>
ODEU_HD void rhs_and_jacobian(double t, const S* z, const S* th, S* f, S (*Df)[Ode::NX]) {
  constexpr int n = Ode::NX;
  using D = GDual<S, 1>;
#pragma unroll 1
  for (int c = 0; c < n; ++c) {
    D Z[n], K[n];
    for (int mm = 0; mm < n; ++mm) { Z[mm].v = z[mm]; Z[mm].d[0] = S(mm == c ? 1.0 : 0.0); }
    Ode::rhs(t, Z, th, K);
    for (int mm = 0; mm < n; ++mm) {
      Df[mm][c] = K[mm].d[0];
      if (c == 0) f[mm] = K[mm].v;
    }
  }
}

constexpr int DIRK_MAX_NEWTON = 500;      // root_find_max_steps of diffrax_solver.py:32

// One DIRK step.  x, th on scalar type S; returns xn, eps (= 0, diffrax_solver.py:128) and the
// step Jacobian J [n][n] = d xn / d x.
template <class Ode, class Tab, class S>
ODEU_HD void dirk_step_generic(double t, double h, const S* x, const S* th, S* xn, S* eps, S (*J)[Ode::NX]) {
  constexpr int n = Ode::NX;
  constexpr int St = Tab::S;
  S Ks[St][n];            // stage derivatives
  S dK[St][n][n];         // their Jacobians w.r.t. x
  S z[n], dZ[n][n];
  for (int i = 0; i < St; ++i) {
    const double ti = t + h * Tab::c(i);
    const double hg = h * Tab::a(i, i);
    S base[n], Bm[n][n];
    for (int mm = 0; mm < n; ++mm) {
      S s = x[mm];
      for (int j = 0; j < i; ++j)
        if (Tab::a(i, j) != 0.0) s = s + Ks[j][mm] * (h * Tab::a(i, j));
      base[mm] = s;
      for (int c = 0; c < n; ++c) {
        S d = S(mm == c ? 1.0 : 0.0);
        for (int j = 0; j < i; ++j)
          if (Tab::a(i, j) != 0.0) d = d + dK[j][mm][c] * (h * Tab::a(i, j));
        Bm[mm][c] = d;
      }
    }
    S f[n], Df[n][n];
    if (hg == 0.0) {                       // explicit first stage of an ESDIRK method
      rhs_and_jacobian<Ode, S>(ti, base, th, f, Df);
      for (int mm = 0; mm < n; ++mm) { z[mm] = base[mm]; for (int c = 0; c < n; ++c) dZ[mm][c] = Bm[mm][c]; }
    } else {
      // predictor: the explicit Euler continuation of the previous stage (any start converges to the
      // same root; the tolerance below, not the start, fixes the result)
      for (int mm = 0; mm < n; ++mm) z[mm] = (i > 0) ? base[mm] + Ks[i - 1][mm] * hg : base[mm];
      int extra = 0;
#pragma unroll 1
      for (int it = 0; it < DIRK_MAX_NEWTON; ++it) {
        rhs_and_jacobian<Ode, S>(ti, z, th, f, Df);
        S A[n][n], R[n][1];
        for (int mm = 0; mm < n; ++mm) {
          R[mm][0] = base[mm] + f[mm] * hg - z[mm];            // -F(z)
          for (int c = 0; c < n; ++c) A[mm][c] = S(mm == c ? 1.0 : 0.0) - Df[mm][c] * hg;
        }
        lu_solve<n, 1, S>(A, R);
        bool conv = true, bad = false;
        for (int mm = 0; mm < n; ++mm) {
          z[mm] = z[mm] + R[mm][0];
          const double dz = fabs(value_of(R[mm][0])), zv = value_of(z[mm]);
          if (!(dz <= 1e-12 * (1.0 + fabs(zv)))) conv = false;
          if (!(fabs(zv) <= 1.7976931348623157e308)) bad = true;              // NaN / inf: give up at once
        }
        if (bad) break;
        if (conv && ++extra == 2) break;   // one more full iteration after convergence (see header)
        if (it + 1 == DIRK_MAX_NEWTON)     // the reference's diffeqsolve fails here; failure is a NaN state
          for (int mm = 0; mm < n; ++mm) z[mm] = z[mm] + std::numeric_limits<double>::quiet_NaN();
      }
      rhs_and_jacobian<Ode, S>(ti, z, th, f, Df);
      S A[n][n];
      for (int mm = 0; mm < n; ++mm)
        for (int c = 0; c < n; ++c) { A[mm][c] = S(mm == c ? 1.0 : 0.0) - Df[mm][c] * hg; dZ[mm][c] = Bm[mm][c]; }
      lu_solve<n, n, S>(A, dZ);            // dZ = (I - h a_ii Df)^-1 (I + h sum a_ij dK_j)
    }
    for (int mm = 0; mm < n; ++mm) {
      Ks[i][mm] = f[mm];
      for (int c = 0; c < n; ++c) {
        S s = S(0.0);
        for (int k = 0; k < n; ++k) s = s + Df[mm][k] * dZ[k][c];
        dK[i][mm][c] = s;
      }
    }
  }
  // stiffly accurate: the last stage is the solution
  for (int mm = 0; mm < n; ++mm) {
    xn[mm] = z[mm];
    eps[mm] = S(0.0);
    for (int c = 0; c < n; ++c) J[mm][c] = dZ[mm][c];
  }
}

}  // namespace odeu
