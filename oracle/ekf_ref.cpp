// ORACLE-B (test infrastructure; never linked into the product, never on the product path).
//
// A plain C++ (std::thread) CPU restatement of the reference's algorithm for the hot path, in the
// reference's own formulation (square-root EKF, every covariance operation a Householder QR of
// stacked transposed factors with LAPACK's dlarfg sign convention).  It is deliberately NOT the
// formulation of the CUDA kernels (full covariance + Cholesky), so agreement between the two is
// an independent check.  Validated against Oracle-A (oracle/ref_torch.py, LAPACK QR through
// torch) by tests/test_oracle_b.py; used as the timed multi-core CPU baseline by bench.py.
//
// Follows, function by function:
//   rk_step           src/solvers/rksolver.py:113-155, compute_node :160-194 (dense ks @ A[i])
//   jmp (tangents)    src/utils.py:72-79 jmp_aux = vmap(jvp(solver)) over the columns of P_sqrt
//   qr_stack          src/utils.py:233-274 sqrt_L_sum_qr / sqrt_L_sum_qr_3 (R^T of economic QR)
//   predict           src/filters/sqrt_ekf.py:92-197 (four noise branches)
//   correct           src/filters/sqrt_ekf.py:337-376 (sign-sensitive zero-gain guard verbatim)
//   nlg               src/utils.py:109-128 negative_log_gaussian_sqrt
//   loop              scripts/run_filter.py:204-222, scripts/run_parameter_estimation.py:771-794
//   ODEs              src/ode/{lorenz,van_der_pol,lotka_volterra,pendulum,lcao,hodgkin_huxley}.py
//   PF predict        src/filters/particle_filter.py:73-118 (noise-free: particle 0 semantics)
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <atomic>
#include <thread>

namespace {

// ---------------------------------------------------------------- forward-mode scalar
struct Dn {  // value + up to KMAX tangents (run-time K)
  static constexpr int KMAX = 16;
  double v;
  double d[KMAX];
};
thread_local int g_K = 0;

inline Dn cst(double c) { Dn r; r.v = c; for (int k = 0; k < g_K; ++k) r.d[k] = 0; return r; }
inline Dn operator+(const Dn& a, const Dn& b) { Dn r; r.v = a.v + b.v; for (int k = 0; k < g_K; ++k) r.d[k] = a.d[k] + b.d[k]; return r; }
inline Dn operator-(const Dn& a, const Dn& b) { Dn r; r.v = a.v - b.v; for (int k = 0; k < g_K; ++k) r.d[k] = a.d[k] - b.d[k]; return r; }
inline Dn operator-(const Dn& a) { Dn r; r.v = -a.v; for (int k = 0; k < g_K; ++k) r.d[k] = -a.d[k]; return r; }
inline Dn operator*(const Dn& a, const Dn& b) { Dn r; r.v = a.v * b.v; for (int k = 0; k < g_K; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k]; return r; }
inline Dn operator/(const Dn& a, const Dn& b) { Dn r; r.v = a.v / b.v; for (int k = 0; k < g_K; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) / b.v; return r; }
inline Dn operator+(const Dn& a, double b) { Dn r = a; r.v = a.v + b; return r; }
inline Dn operator+(double a, const Dn& b) { Dn r = b; r.v = a + b.v; return r; }
inline Dn operator-(const Dn& a, double b) { Dn r = a; r.v = a.v - b; return r; }
inline Dn operator-(double a, const Dn& b) { Dn r; r.v = a - b.v; for (int k = 0; k < g_K; ++k) r.d[k] = -b.d[k]; return r; }
inline Dn operator*(const Dn& a, double b) { Dn r; r.v = a.v * b; for (int k = 0; k < g_K; ++k) r.d[k] = a.d[k] * b; return r; }
inline Dn operator*(double a, const Dn& b) { return b * a; }
inline Dn operator/(const Dn& a, double b) { Dn r; r.v = a.v / b; for (int k = 0; k < g_K; ++k) r.d[k] = a.d[k] / b; return r; }
inline Dn operator/(double a, const Dn& b) { Dn r; r.v = a / b.v; for (int k = 0; k < g_K; ++k) r.d[k] = -r.v * b.d[k] / b.v; return r; }
inline Dn ex(const Dn& a) { Dn r; r.v = std::exp(a.v); for (int k = 0; k < g_K; ++k) r.d[k] = r.v * a.d[k]; return r; }
inline Dn sn(const Dn& a) { Dn r; r.v = std::sin(a.v); const double c = std::cos(a.v); for (int k = 0; k < g_K; ++k) r.d[k] = c * a.d[k]; return r; }
inline double ex(double a) { return std::exp(a); }
inline double sn(double a) { return std::sin(a); }
inline double cst_like(double, double c) { return c; }
inline Dn cst_like(const Dn&, double c) { return cst(c); }

// ---------------------------------------------------------------- ODE right-hand sides
enum { ODE_LORENZ = 0, ODE_VDP, ODE_LV, ODE_PENDULUM, ODE_LCAO, ODE_HH, ODE_MULTI_HH };

template <class T> T hh_single(int model, double t, const T* x, const double* p /*15, stride*/, int st, T* dx) {
  // p order: C, A, g_Na, E_Na, g_K, E_K, g_leak, E_leak, V_T, g_M, tau_max, g_L, E_Ca, g_T, V_x
  const double C = p[0 * st], A = p[1 * st], gNa = p[2 * st], ENa = p[3 * st], gK = p[4 * st],
               EK = p[5 * st], gl = p[6 * st], El = p[7 * st], VT = p[8 * st], gM = p[9 * st],
               tmax = p[10 * st], gL = p[11 * st], ECa = p[12 * st], gT = p[13 * st], Vx = p[14 * st];
  const T V = x[0];
  // src/ode/hodgkin_huxley.py:12-27
  const T a_m = -0.32 * (V - VT - 13.0) / (ex(-(V - VT - 13.0) / 4.0) - 1.0);
  const T b_m = 0.28 * (V - VT - 40.0) / (ex((V - VT - 40.0) / 5.0) - 1.0);
  const T a_n = -0.032 * (V - VT - 15.0) / (ex(-(V - VT - 15.0) / 5.0) - 1.0);
  const T b_n = 0.5 * ex(-(V - VT - 10.0) / 40.0);
  const T a_h = 0.128 * ex(-(V - VT - 17.0) / 18.0);
  const T b_h = 4.0 / (1.0 + ex(-(V - VT - 40.0) / 5.0));
  dx[1] = a_m * (1.0 - x[1]) - b_m * x[1];
  dx[2] = a_h * (1.0 - x[2]) - b_h * x[2];
  dx[3] = a_n * (1.0 - x[3]) - b_n * x[3];
  T I = gNa * (x[1] * x[1] * x[1]) * x[2] * (ENa - V);
  { T n2 = x[3] * x[3]; I = I + gK * (n2 * n2) * (EK - V); }
  I = I + gl * (El - V);
  if (model != 4) {
    const T p_inf = 1.0 / (1.0 + ex(-(V + 35.0) / 10.0));
    const T tau_p = tmax / (3.3 * ex((V + 35.0) / 20.0) + ex(-(V + 35.0) / 20.0));
    const T a_q = 0.055 * (-27.0 - V) / (ex((-27.0 - V) / 3.8) - 1.0);
    const T b_q = 0.94 * ex((-75.0 - V) / 17.0);
    const T a_r = 0.000457 * ex((-13.0 - V) / 50.0);
    const T b_r = 0.0065 / (ex((-15.0 - V) / 28.0) + 1.0);
    dx[4] = (p_inf - x[4]) / tau_p;
    dx[5] = a_q * (1.0 - x[5]) - b_q * x[5];
    dx[6] = a_r * (1.0 - x[6]) - b_r * x[6];
    I = I + gM * x[4] * (EK - V);
    I = I + gL * (x[5] * x[5]) * x[6] * (ECa - V);
  }
  if (model == 0) {
    const T tau_u = (30.8 + (211.4 + ex((V + Vx + 113.2) / 5.0))) / (3.7 * (1.0 + ex((V + Vx + 84.0) / 3.2)));
    const T u_inf = 1.0 / (1.0 + ex((V + Vx + 81.0) / 4.0));
    const T s_inf = 1.0 / (1.0 + ex(-(V + Vx + 57.0) / 6.2));
    dx[7] = (u_inf - x[7]) / tau_u;
    I = I + gT * (s_inf * s_inf) * x[7] * (ECa - V);
  }
  const double I_in = (t >= 10.0 && t <= 90.0) ? 210.0 * 1e-6 : 0.0;
  dx[0] = (I + I_in / A) / C;
  return dx[0];
}

inline int hh_dim(int model) { return model == 0 ? 8 : model == 1 ? 7 : 4; }

struct OdeSpec { int id, variant, nc, n, p; };

template <class T> void rhs(const OdeSpec& o, double t, const T* x, const double* th, T* dx) {
  switch (o.id) {
    case ODE_LORENZ:  // src/ode/lorenz.py:46-52 (sigma, beta, rho)
      dx[0] = th[0] * (x[1] - x[0]);
      dx[1] = x[0] * (th[2] - x[2]) - x[1];
      dx[2] = x[0] * x[1] - th[1] * x[2];
      break;
    case ODE_VDP:  // src/ode/van_der_pol.py:38-44
      dx[0] = x[1];
      dx[1] = th[0] * (1.0 - x[0] * x[0]) * x[1] - x[0];
      break;
    case ODE_LV:  // src/ode/lotka_volterra.py:47-52
      dx[0] = th[0] * x[0] - th[1] * x[0] * x[1];
      dx[1] = (-th[2]) * x[1] + th[3] * x[0] * x[1];
      break;
    case ODE_PENDULUM:  // src/ode/pendulum.py:38-44
      dx[0] = x[1];
      dx[1] = (-9.81 / th[0]) * sn(x[0]);
      break;
    case ODE_LCAO: {  // src/ode/lcao.py:51-61
      const int D = o.variant;
      for (int i = 0; i < D; ++i) {
        dx[i] = x[D + i];
        dx[D + i] = (-th[0]) * x[i] - th[1] * (x[i] * x[i] * x[i]) - th[2] * x[D - 1 - i];
      }
    } break;
    case ODE_HH: hh_single<T>(o.variant, t, x, th, 1, dx); break;
    case ODE_MULTI_HH: {  // src/ode/hodgkin_huxley.py:358-401
      const int NC = o.nc, D = hh_dim(o.variant);
      const double* cc = th;
      const double Cm = th[NC - 1];
      const double* per = th + NC;
      for (int c = 0; c < NC; ++c) {
        double loc[15];
        loc[0] = Cm;
        for (int k = 1; k < 15; ++k) loc[k] = per[(k - 1) * NC + c];
        hh_single<T>(o.variant, t, x + c * D, loc, 1, dx + c * D);
      }
      // V_coupled = G @ V with G tridiagonal, G_cc = -(cc[c-1] + cc[c])
      for (int c = 0; c < NC; ++c) {
        double gd = 0.0;
        if (c > 0) gd += -cc[c - 1];
        if (c + 1 < NC) gd += -cc[c];
        T acc = gd * x[c * D];
        if (c > 0) acc = acc + cc[c - 1] * x[(c - 1) * D];
        if (c + 1 < NC) acc = acc + cc[c] * x[(c + 1) * D];
        dx[c * D] = dx[c * D] + acc / Cm;
      }
    } break;
  }
}

// ---------------------------------------------------------------- tableaux (dense, zeros kept)
struct Tableau { int S; double A[8][8]; double b[2][8]; double c[8]; };

Tableau make_tableau(int id) {
  Tableau t; std::memset(&t, 0, sizeof(t));
  if (id == 0) {  // RKF45 src/solvers/rkf45.py:10-34
    t.S = 6;
    const double A[6][6] = {{0}, {1.0 / 4}, {3.0 / 32, 9.0 / 32}, {1932.0 / 2197, -7200.0 / 2197, 7296.0 / 2197},
                            {439.0 / 216, -8.0, 3680.0 / 513, -845.0 / 4104},
                            {-8.0 / 27, 2.0, -3544.0 / 2565, 1859.0 / 4104, -11.0 / 40}};
    const double b[2][6] = {{16.0 / 135, 0, 6656.0 / 12825, 28561.0 / 56430, -9.0 / 50, 2.0 / 55},
                            {25.0 / 216, 0, 1408.0 / 2565, 2197.0 / 4104, -1.0 / 5, 0}};
    const double c[6] = {0, 1.0 / 4, 3.0 / 8, 12.0 / 13, 1, 1.0 / 2};
    for (int i = 0; i < 6; ++i) { t.c[i] = c[i]; t.b[0][i] = b[0][i]; t.b[1][i] = b[1][i]; for (int j = 0; j < 6; ++j) t.A[i][j] = A[i][j]; }
  } else if (id == 1) {  // Dopri65 src/solvers/dopri65.py:10-72
    t.S = 8;
    const double A[8][8] = {{0}, {1.0 / 10}, {-2.0 / 81, 20.0 / 81}, {615.0 / 1372, -270.0 / 343, 1053.0 / 1372},
                            {3243.0 / 5500, -54.0 / 55, 50949.0 / 71500, 4998.0 / 17875},
                            {-26492.0 / 37125, 72.0 / 55, 2808.0 / 23375, -24206.0 / 37125, 338.0 / 459},
                            {5561.0 / 2376, -35.0 / 11, -24117.0 / 31603, 899983.0 / 200772, -5225.0 / 1836, 3925.0 / 4056},
                            {465467.0 / 266112, -2945.0 / 1232, -5610201.0 / 14158144, 10513573.0 / 3212352,
                             -424325.0 / 205632, 376225.0 / 454272}};
    const double b[2][8] = {{821.0 / 10800, 0, 19683.0 / 71825, 175273.0 / 912600, 395.0 / 3672, 785.0 / 2704, 3.0 / 50, 0},
                            {61.0 / 864, 0, 98415.0 / 321776, 16807.0 / 146016, 1375.0 / 7344, 1375.0 / 5408, -37.0 / 1120, 1.0 / 10}};
    const double c[8] = {0, 1.0 / 10, 2.0 / 9, 3.0 / 7, 3.0 / 5, 4.0 / 5, 1.0, 1.0};
    for (int i = 0; i < 8; ++i) { t.c[i] = c[i]; t.b[0][i] = b[0][i]; t.b[1][i] = b[1][i]; for (int j = 0; j < 8; ++j) t.A[i][j] = A[i][j]; }
  } else if (id == 2) {  // BS32 src/solvers/bs32.py:10-32
    t.S = 4;
    const double A[4][4] = {{0}, {1.0 / 2}, {0, 3.0 / 4}, {2.0 / 9, 1.0 / 3, 4.0 / 9}};
    const double b[2][4] = {{7.0 / 24, 1.0 / 4, 1.0 / 3, 1.0 / 8}, {2.0 / 9, 1.0 / 3, 4.0 / 9, 0}};
    const double c[4] = {0, 1.0 / 2, 3.0 / 4, 1.0};
    for (int i = 0; i < 4; ++i) { t.c[i] = c[i]; t.b[0][i] = b[0][i]; t.b[1][i] = b[1][i]; for (int j = 0; j < 4; ++j) t.A[i][j] = A[i][j]; }
  } else {  // HeunEuler src/solvers/heun_euler.py:10-30 (b[1] = [0.5, 0] verbatim)
    t.S = 2;
    t.A[1][0] = 1.0; t.b[0][0] = 0.5; t.b[0][1] = 0.5; t.b[1][0] = 0.5; t.b[1][1] = 0.0; t.c[1] = 1.0;
  }
  return t;
}

constexpr int NMAX = 16;

// One RK step on scalar kind T (double or Dn).  x_next[r] = x + h * (ks @ b[r]).
template <class T>
void rk_step(const OdeSpec& o, const Tableau& tb, double h, double t, const T* x, const double* th,
             T* x_next1, T* x_next0) {
  const int n = o.n, S = tb.S;
  T ks[8][NMAX];
  for (int i = 0; i < S; ++i) {
    T xi[NMAX];
    for (int m = 0; m < n; ++m) {
      T s = cst_like(x[m], 0.0);
      for (int j = 0; j < S; ++j) {      // dense `ks @ A[idx]` incl. zeros (rksolver.py:193)
        if (j < i) s = s + ks[j][m] * tb.A[i][j];
      }
      xi[m] = x[m] + h * s;
    }
    rhs<T>(o, t + h * tb.c[i], xi, th, ks[i]);
  }
  for (int m = 0; m < n; ++m) {
    T s0 = cst_like(x[m], 0.0), s1 = cst_like(x[m], 0.0);
    for (int j = 0; j < S; ++j) { s0 = s0 + ks[j][m] * tb.b[0][j]; s1 = s1 + ks[j][m] * tb.b[1][j]; }
    x_next0[m] = x[m] + h * s0;
    x_next1[m] = x[m] + h * s1;
  }
}

// ---------------------------------------------------------------- Householder QR, R only
// A is m x n row-major (m >= n).  On return the upper triangle of the first n rows holds R with
// LAPACK dgeqr2/dlarfg signs: beta = -sign(alpha) * norm (Fortran SIGN: the sign BIT of alpha);
// a sub-column that is STRUCTURALLY zero gives H = I, beta = alpha (dlarfg, xnorm == 0).
//
// Fragile columns.  The sign of R_jj is -sign(alpha_j) when the sub-column is non-zero and
// +sign(alpha_j) when it is exactly zero.  When the columns of the stacked matrix are (nearly)
// collinear - a covariance that has collapsed onto fewer directions than n, e.g. Van der Pol with
// observation noise >> process noise - the sub-column after the earlier reflectors is pure
// cancellation noise, and whether that noise is exactly 0.0 or 1e-22 depends on the BLAS build
// (FMA or not, accumulation order): the reference's OWN sign, and with it the guard
// `all(S_sqrt < 1e-16)`, is then decided by rounding.  Measured against LAPACK (torch.linalg.qr,
// tests/test_guard_reference.py): the generic outcome is "noise != 0".  This routine therefore
//   (a) treats a sub-column that CANCELLED to exactly zero (entries that received non-zero update
//       terms) like noise: beta = -sign(alpha) |alpha|, i.e. the reflector flips row j, and
//   (b) counts every column whose sub-column norm is at the rounding level of the terms that
//       produced it (<= 1e3 ulp) in *fragile, so a parity run can tell "sign decided by
//       arithmetic" from "sign decided by rounding".
void householder_R(double* A, int m, int n, long long* fragile = nullptr) {
  // touched[i][k]: entry received a non-zero update term (tells "cancelled to zero" from "never set");
  // Mg (only when counting fragile columns): magnitude of the terms each entry was built from
  std::vector<unsigned char> touched((size_t)m * n, 0);
  std::vector<double> Mg;
  if (fragile) { Mg.resize((size_t)m * n); for (int i = 0; i < m * n; ++i) Mg[i] = std::fabs(A[i]); }
  for (int j = 0; j < n; ++j) {
    double xnorm2 = 0.0, mg2 = 0.0;
    bool any_touched = false;
    for (int i = j + 1; i < m; ++i) {
      xnorm2 += A[i * n + j] * A[i * n + j];
      any_touched = any_touched || touched[i * n + j];
      if (fragile) mg2 += Mg[i * n + j] * Mg[i * n + j];
    }
    const double alpha = A[j * n + j];
    if (fragile && mg2 > 0.0 && xnorm2 <= (1e3 * 1.1e-16) * (1e3 * 1.1e-16) * mg2) ++*fragile;
    if (xnorm2 == 0.0) {
      if (!any_touched) continue;              // structurally zero: dlarfg gives H = I, beta = alpha
      A[j * n + j] = -alpha;                   // cancelled to zero: generic LAPACK outcome (see above)
      for (int k = j + 1; k < n; ++k) A[j * n + k] = -A[j * n + k];
      continue;
    }
    const double norm = std::sqrt(alpha * alpha + xnorm2);
    const double beta = std::signbit(alpha) ? norm : -norm;
    const double tau = (beta - alpha) / beta;
    const double scal = 1.0 / (alpha - beta);
    // v = [1, A[j+1:, j] * scal]; apply H = I - tau v v^T to the trailing columns
    for (int i = j + 1; i < m; ++i) A[i * n + j] *= scal;
    A[j * n + j] = beta;
    for (int k = j + 1; k < n; ++k) {
      double w = A[j * n + k], wm = fragile ? Mg[j * n + k] : 0.0;
      for (int i = j + 1; i < m; ++i) {
        w += A[i * n + j] * A[i * n + k];
        if (fragile) wm += std::fabs(A[i * n + j]) * Mg[i * n + k];
      }
      w *= tau; wm *= std::fabs(tau);
      A[j * n + k] -= w;
      if (fragile) Mg[j * n + k] += wm;
      for (int i = j + 1; i < m; ++i) {
        const double upd = w * A[i * n + j];
        A[i * n + k] -= upd;
        if (upd != 0.0) touched[i * n + k] = 1;
        if (fragile) Mg[i * n + k] += wm * std::fabs(A[i * n + j]);
      }
    }
  }
}

// out (n x n, lower) = R^T of qr([blk0^T; blk1^T; ...]); each block is rows x n? No: each block
// b_k is an [n x c_k] factor; the stacked matrix has sum(c_k) rows and n columns.
thread_local long long g_fragile = 0;     // fragile columns seen by this thread (see householder_R)
bool g_count_fragile = false;             // set per run (diagnostic; off for the timed CPU baseline)
void qr_stack(int n, int nblk, const double* const* blk, const int* cols, double* out) {
  int m = 0;
  for (int k = 0; k < nblk; ++k) m += cols[k];
  std::vector<double> A((size_t)m * n);
  int r0 = 0;
  for (int k = 0; k < nblk; ++k) {
    for (int c = 0; c < cols[k]; ++c)
      for (int i = 0; i < n; ++i) A[(size_t)(r0 + c) * n + i] = blk[k][i * cols[k] + c];  // transpose
    r0 += cols[k];
  }
  householder_R(A.data(), m, n, g_count_fragile ? &g_fragile : nullptr);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) out[i * n + j] = (j <= i && j < m) ? A[(size_t)j * n + i] : 0.0;
}

struct Cfg {
  OdeSpec ode; Tableau tb; double h; int cov_fn; double scale; int disable;
  int L; int guard_intended;
};

// predict (sqrt_ekf.py:92-197): returns new x, eps, P_sqrt (n x n, maybe non-triangular)
void predict(const Cfg& c, double t, double* x, double* eps, double* Ps, const double* th,
             const double* Q_sqrt, double gamma_sqrt, bool qany) {
  const int n = c.ode.n;
  g_K = n;
  Dn X[NMAX], X1[NMAX], X0[NMAX];
  for (int i = 0; i < n; ++i) { X[i].v = x[i]; for (int k = 0; k < n; ++k) X[i].d[k] = Ps[i * n + k]; }
  rk_step<Dn>(c.ode, c.tb, c.h, t, X, th, X1, X0);
  double T[NMAX * NMAX];
  for (int i = 0; i < n; ++i) {
    x[i] = X1[i].v;
    eps[i] = std::fabs(X0[i].v - X1[i].v);
    for (int k = 0; k < n; ++k) T[i * n + k] = X1[i].d[k];
  }
  double B1[NMAX * NMAX], B2[NMAX * NMAX];
  if (c.disable) {
    if (qany) {  // sqrt_L_sum_qr(T, gamma_sqrt * Q_sqrt)
      for (int i = 0; i < n * n; ++i) B1[i] = gamma_sqrt * Q_sqrt[i];
      const double* blk[2] = {T, B1}; const int cols[2] = {n, n};
      qr_stack(n, 2, blk, cols, Ps);
    } else {
      std::memcpy(Ps, T, sizeof(double) * n * n);
    }
  } else if (qany) {  // sqrt_L_sum_qr_3(gamma_sqrt*Q_sqrt, diag(eps), T)
    for (int i = 0; i < n * n; ++i) { B1[i] = gamma_sqrt * Q_sqrt[i]; B2[i] = 0.0; }
    for (int i = 0; i < n; ++i) B2[i * n + i] = eps[i];
    const double* blk[3] = {B1, B2, T}; const int cols[3] = {n, n, n};
    qr_stack(n, 3, blk, cols, Ps);
  } else {
    for (int i = 0; i < n * n; ++i) B1[i] = 0.0;
    if (c.cov_fn == 0) {
      for (int i = 0; i < n; ++i) B1[i * n + i] = c.scale * eps[i];
    } else if (c.cov_fn == 1) {
      double ss = 0.0;
      for (int i = 0; i < n; ++i) ss += (c.scale * eps[i]) * (c.scale * eps[i]);
      const double nrm = std::sqrt(ss);
      for (int i = 0; i < n; ++i)
        for (int k = 0; k < n; ++k) B1[i * n + k] = (c.scale * eps[i]) * (c.scale * eps[k]) / nrm;
    } else {
      for (int i = 0; i < n; ++i) B1[i * n + i] = c.scale;
    }
    const double* blk[2] = {T, B1}; const int cols[2] = {n, n};
    qr_stack(n, 2, blk, cols, Ps);
  }
}

// correct (sqrt_ekf.py:337-376) + nlg (utils.py:109-128).  Returns nlg; sets *mismatch when the
// verbatim and intended guards disagree.
double correct(const Cfg& c, double* x, double* Ps, const double* y, const double* H,
               const double* R_sqrt, double* yhat, double* S_sqrt, int* mismatch, int* fired) {
  const int n = c.ode.n, L = c.L;
  double HP[NMAX * NMAX];  // [L x n]
  double d[NMAX];
  for (int l = 0; l < L; ++l) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) s += H[l * n + j] * x[j];
    yhat[l] = s;
    d[l] = y[l] - s;
    for (int k = 0; k < n; ++k) {
      double a = 0.0;
      for (int j = 0; j < n; ++j) a += H[l * n + j] * Ps[j * n + k];
      HP[l * n + k] = a;
    }
  }
  { const double* blk[2] = {HP, R_sqrt}; const int cols[2] = {n, L}; qr_stack(L, 2, blk, cols, S_sqrt); }
  bool g_ref = true, g_int = true;
  for (int i = 0; i < L * L; ++i) { g_ref = g_ref && (S_sqrt[i] < 1e-16); g_int = g_int && (std::fabs(S_sqrt[i]) < 1e-16); }
  *mismatch += (g_ref != g_int);
  const bool zeroK = c.guard_intended ? g_int : g_ref;
  *fired += zeroK;
  // K = (cho_solve((S_sqrt, lower), H) @ P_sqrt @ P_sqrt^T)^T
  double K[NMAX * NMAX];  // [n x L]
  if (zeroK) {
    for (int i = 0; i < n * L; ++i) K[i] = 0.0;
  } else {
    double Z[NMAX * NMAX];  // S^-1 H  [L x n]
    for (int j = 0; j < n; ++j) {
      double w[NMAX];
      for (int l = 0; l < L; ++l) {
        double s = H[l * n + j];
        for (int k = 0; k < l; ++k) s -= S_sqrt[l * L + k] * w[k];
        w[l] = s / S_sqrt[l * L + l];
      }
      for (int l = L - 1; l >= 0; --l) {
        double s = w[l];
        for (int k = l + 1; k < L; ++k) s -= S_sqrt[k * L + l] * Z[k * n + j];
        Z[l * n + j] = s / S_sqrt[l * L + l];
      }
    }
    double ZP[NMAX * NMAX];  // Z @ P_sqrt  [L x n]
    for (int l = 0; l < L; ++l)
      for (int k = 0; k < n; ++k) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) s += Z[l * n + j] * Ps[j * n + k];
        ZP[l * n + k] = s;
      }
    for (int l = 0; l < L; ++l)
      for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += ZP[l * n + k] * Ps[i * n + k];
        K[i * L + l] = s;
      }
  }
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int l = 0; l < L; ++l) s += K[i * L + l] * d[l];
    x[i] += s;
  }
  // Joseph: sqrt_L_sum_qr((I - K H) P_sqrt, K R_sqrt)
  double A[NMAX * NMAX], AP[NMAX * NMAX], KR[NMAX * NMAX];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int l = 0; l < L; ++l) s -= K[i * L + l] * H[l * n + j];
      A[i * n + j] = s;
    }
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < n; ++k) {
      double s = 0.0;
      for (int j = 0; j < n; ++j) s += A[i * n + j] * Ps[j * n + k];
      AP[i * n + k] = s;
    }
  for (int i = 0; i < n; ++i)
    for (int m = 0; m < L; ++m) {
      double s = 0.0;
      for (int l = 0; l < L; ++l) s += K[i * L + l] * R_sqrt[l * L + m];
      KR[i * L + m] = s;
    }
  { const double* blk[2] = {AP, KR}; const int cols[2] = {n, L}; qr_stack(n, 2, blk, cols, Ps); }
  // nlg
  double z[NMAX], quad = 0.0, logdet = 0.0;
  for (int i = 0; i < L; ++i) {
    double s = d[i];
    for (int k = 0; k < i; ++k) s -= S_sqrt[i * L + k] * z[k];
    z[i] = s / S_sqrt[i * L + i];
    quad += z[i] * z[i];
    logdet += std::log(std::fabs(S_sqrt[i * L + i]));
  }
  return 0.5 * quad + 0.5 * L * std::log(2.0 * M_PI) + logdet;
}

}  // namespace

extern "C" {

// All arrays HOST, trajectory-major ("reference orientation"):
//   x0 [B][n], theta [B][p] or shared [p] (theta_per_traj), ys [T_obs][L] or [T_obs][B][L],
//   outputs xT [B][n], PT [B][n][n] (= P_sqrt P_sqrt^T), nll [B]; optional trajectory
//   out_x/out_eps [Ts][B][n], out_P [Ts][B][n][n], out_yhat [Ts][B][L], out_S [Ts][B][L][L], out_t [Ts]
// Returns the number of steps on which the verbatim and intended zero-gain guards disagreed.
long long oracle_ekf_run(int ode_id, int variant, int nc, int n, int p, int solver_id, double h,
                         int cov_fn, double scale, int disable, long long B, long long T, double t0,
                         const double* x0, const double* P0_sqrt, const double* theta,
                         int theta_per_traj, const double* Q_sqrt, double gamma_sqrt, int L,
                         const double* H, const double* R_sqrt, const double* ys, int ys_per_traj,
                         const unsigned char* flags, const long long* ymap, long long save_interval,
                         int guard_intended, int nthreads, double* xT, double* PT, double* nll,
                         double* out_t, double* out_x, double* out_eps, double* out_P,
                         double* out_yhat, double* out_S, long long* fired_out, long long* fragile_out) {
  Cfg c;
  c.ode = {ode_id, variant, nc, n, p};
  c.tb = make_tableau(solver_id);
  c.h = h; c.cov_fn = cov_fn; c.scale = scale; c.disable = disable; c.L = L;
  c.guard_intended = guard_intended;
  bool qany = false;
  std::vector<double> Qz(n * n, 0.0);
  if (Q_sqrt) for (int i = 0; i < n * n; ++i) { Qz[i] = Q_sqrt[i]; qany = qany || (Q_sqrt[i] >= 1e-16); }
  std::atomic<long long> mismatch_total{0}, fired_total{0}, fragile_total{0};
  g_count_fragile = fragile_out != nullptr;
  auto run_one = [&](long long b) {
    g_K = n;
    const long long fragile0 = g_fragile;
    double x[NMAX], eps[NMAX], Ps[NMAX * NMAX], yhat[NMAX], S_sqrt[NMAX * NMAX];
    for (int i = 0; i < n; ++i) { x[i] = x0[b * n + i]; eps[i] = 0.0; yhat[i] = 0.0; }
    for (int i = 0; i < n * n; ++i) Ps[i] = P0_sqrt[i];
    for (int i = 0; i < L * L; ++i) S_sqrt[i] = 0.0;
    const double* th = theta_per_traj ? theta + b * p : theta;
    double t = t0, acc = 0.0;
    int mismatch = 0, fired = 0;
    auto save = [&](long long slot) {
      if (out_x) for (int i = 0; i < n; ++i) out_x[(slot * B + b) * n + i] = x[i];
      if (out_eps) for (int i = 0; i < n; ++i) out_eps[(slot * B + b) * n + i] = eps[i];
      if (out_P)
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) s += Ps[i * n + k] * Ps[j * n + k];
            out_P[((slot * B + b) * n + i) * n + j] = s;
          }
      if (out_yhat) for (int l = 0; l < L; ++l) out_yhat[(slot * B + b) * L + l] = yhat[l];
      if (out_S)
        for (int i = 0; i < L; ++i)
          for (int j = 0; j < L; ++j) {
            double s = 0.0;
            for (int k = 0; k < L; ++k) s += S_sqrt[i * L + k] * S_sqrt[j * L + k];
            out_S[((slot * B + b) * L + i) * L + j] = s;
          }
      if (out_t && b == 0) out_t[slot] = t;
    };
    if (save_interval > 0) save(0);
    for (long long step = 0; step < T; ++step) {
      predict(c, t, x, eps, Ps, th, Qz.data(), gamma_sqrt, qany);
      t = t + h;
      if (L > 0 && flags[step]) {
        const long long oi = ymap[step];
        const double* y = ys_per_traj ? ys + (oi * B + b) * L : ys + oi * L;
        acc += correct(c, x, Ps, y, H, R_sqrt, yhat, S_sqrt, &mismatch, &fired);
      }
      if (save_interval > 0 && (step + 1) % save_interval == 0) save((step + 1) / save_interval);
    }
    if (xT) for (int i = 0; i < n; ++i) xT[b * n + i] = x[i];
    if (PT)
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          double s = 0.0;
          for (int k = 0; k < n; ++k) s += Ps[i * n + k] * Ps[j * n + k];
          PT[(b * n + i) * n + j] = s;
        }
    if (nll) nll[b] = acc;
    mismatch_total += mismatch;
    fired_total += fired;
    fragile_total += g_fragile - fragile0;
  };
  int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if ((long long)nt > B) nt = (int)B;
  if (nt == 1) {
    for (long long b = 0; b < B; ++b) run_one(b);
  } else {
    std::vector<std::thread> pool;
    for (int w = 0; w < nt; ++w)
      pool.emplace_back([&, w]() {
        const long long lo = B * w / nt, hi = B * (w + 1) / nt;   // contiguous shard per core
        for (long long b = lo; b < hi; ++b) run_one(b);
      });
    for (auto& th : pool) th.join();
  }
  if (fired_out) *fired_out = fired_total.load();
  if (fragile_out) *fragile_out = fragile_total.load();   // QR columns whose sign was decided by rounding
  return mismatch_total.load();
}

// Plain RK trajectory (noise-free particle / data synthesis): out_x [T+1][n], out_eps [T+1][n].
void oracle_rk_run(int ode_id, int variant, int nc, int n, int p, int solver_id, double h, long long T,
                   double t0, const double* x0, const double* theta, double* out_x, double* out_eps) {
  OdeSpec o = {ode_id, variant, nc, n, p};
  Tableau tb = make_tableau(solver_id);
  double x[NMAX], t = t0;
  for (int i = 0; i < n; ++i) { x[i] = x0[i]; out_x[i] = x0[i]; out_eps[i] = 0.0; }
  for (long long s = 0; s < T; ++s) {
    double x1[NMAX], xz[NMAX];
    rk_step<double>(o, tb, h, t, x, theta, x1, xz);
    for (int i = 0; i < n; ++i) {
      out_eps[(s + 1) * n + i] = std::fabs(xz[i] - x1[i]);
      x[i] = x1[i];
      out_x[(s + 1) * n + i] = x[i];
    }
    t = t + h;
  }
}

int oracle_num_threads(void) {
  const int nt = (int)std::thread::hardware_concurrency();
  return nt > 0 ? nt : 1;
}

}  // extern "C"
