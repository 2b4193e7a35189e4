// Column-parallel cooperative EKF kernel for medium state dimensions (5 <= n <= 16: the
// Hodgkin-Huxley family; BASELINE config 3).
//
// With B = 4,096 parameter sets a thread-per-trajectory mapping leaves the GPU almost empty
// (128 warps on 148 SMs) and every thread serialises n tangent passes.  Here a CTA owns TB
// trajectories (or (trajectory, parameter-direction) pairs for the gradient) and n*TB threads:
// thread (tl, c) propagates tangent COLUMN c of trajectory tl through the RK step, then the
// threads of a trajectory combine their columns through shared memory:
//
//   phase 1  RK step with one tangent column (identity seed e_c)   -> J[:, c]  -> smem
//   phase 2  M[:, c] = J P[:, c]            (P is kept distributed: thread c owns column c)
//   phase 3  P+[:, c] = M J[c, :]^T + Q[:, c]
//   phase 4  measurement update, column-wise: PHt row c = P[:, c] . H rows is local by symmetry,
//            S / Cholesky / NLL are recomputed per thread (L <= n is small), gain row c, and the
//            Joseph form evaluated as AP - G K^T with G = PHt - K S (its rounding-level residual)
//
// The scalar type S is `double` (NLL only) or GDual<double,1> (NLL + one parameter direction per
// thread column group): the same statements produce value and derivative.  All shared arrays are
// laid out [..][TB] with the trajectory index fastest (bank-conflict free); threads of one warp
// share c, so the code is warp-uniform.
//
// The per-thread body is written as a struct with one method per phase so the identical source
// is replayed sequentially on the host by the test-only emulation (tests/host_emu.cu).
#pragma once
#include "ekf_grad.cuh"

namespace odeu {

template <class S> struct scalar_ops;
template <> struct scalar_ops<double> {
  ODEU_HD static double val(double a) { return a; }
};
template <int K> struct scalar_ops<GDual<double, K>> {
  ODEU_HD static double val(const GDual<double, K>& a) { return a.v; }
};

template <class Ode, class Tab, class S, int TB>
struct CoopThread {
  static constexpr int n = Ode::NX;
  static constexpr int NP = Ode::NP;
  using Args = GradArgs<Ode::NX, Ode::NP>;

  // ---- per-thread state
  int tl, c;            // trajectory slot in the CTA, owned column
  long long b;          // trajectory index
  int chunk;            // parameter direction (gradient) or 0
  bool active;
  S x[n], eps[n], Pcol[n], th[NP];
  S xn[n];
  S nll;
  double t;
  // scratch that lives across phases
  S ph[n];              // PHt row c            (L used)
  S Krow[n];            // gain row c           (L used)
  S dvec[n];            // innovation           (L used)

  // shared memory views: A = J [n][n][TB];  Bm = M [n][n][TB], later PHt/K/G [3][n][L][TB]
  ODEU_HD static S& at2(S* base, int i, int j, int tl) { return base[((long long)i * n + j) * TB + tl]; }
  ODEU_HD static S& at3(S* base, int which, int i, int l, int L, int tl) {
    return base[(((long long)which * n + i) * L + l) * TB + tl];
  }

  ODEU_HD void init(const Args& a, long long unit, int tl_, int c_) {
    tl = tl_; c = c_;
    const long long total = a.B * (a.p_opt > 0 ? a.p_opt : 1);
    active = unit < total;
    b = active ? unit % a.B : 0;
    chunk = active ? (int)(unit / a.B) : 0;
    for (int i = 0; i < n; ++i) {
      x[i] = S(a.x0[i * a.B + b]);
      eps[i] = S(0.0);
      Pcol[i] = S(a.P0s[i * n + c]);
    }
    for (int k = 0; k < NP; ++k) th[k] = S(a.theta ? a.theta[k * a.B + b] : a.theta_shared[k]);
    seed_direction(a);
    nll = S(0.0);
    t = a.t0;
  }
  // (gamma_sqrt w_c)^2 of the per-trajectory diagonal Q (parameter_sensitivity); re-read every step
  // (an L1/L2 hit) instead of living in registers for the whole run
  ODEU_HD S q_own(const Args& a) const {
    S wv = S(a.qdiag[c * a.B + b]);
    if constexpr (!std::is_same<S, double>::value) {
      if (chunk < a.p_opt && a.qdiag_tan) wv.d[0] = a.qdiag_tan[((long long)chunk * n + c) * a.B + b];
    }
    wv = wv * a.q_gamma;
    return wv * wv;
  }
  // parameter direction of this thread (gradient instantiation only)
  ODEU_HD void seed_direction(const Args& a) {
    if constexpr (!std::is_same<S, double>::value) {
      if (chunk < a.p_opt) {
        th[a.idx[chunk]].d[0] = 1.0;
        if (a.x0_tan)
          for (int i = 0; i < n; ++i) x[i].d[0] = a.x0_tan[((long long)chunk * n + i) * a.B + b];
      }
    }
  }

  // phase 1: RK step with tangent column c -> J[:, c]
  ODEU_HD void phase_rk(const Args& a, S* Asm) {
    S Jc[n];
    rk_step_rolled<Ode, Tab::S>(a, t, x, th, c, xn, eps, Jc);
    for (int i = 0; i < n; ++i) at2(Asm, i, c, tl) = Jc[i];
  }
  // phase 2: M[:, c] = J P[:, c]
  ODEU_HD void phase_m(S* Asm, S* Bsm) {
    for (int i = 0; i < n; ++i) {
      S s = at2(Asm, i, 0, tl) * Pcol[0];
      for (int j = 1; j < n; ++j) s = s + at2(Asm, i, j, tl) * Pcol[j];
      at2(Bsm, i, c, tl) = s;
    }
  }
  // phase 3: P+[:, c] = M J[c, :]^T + Q[:, c];  x <- x_next;  t <- t + h
  ODEU_HD void phase_p(const Args& a, S* Asm, S* Bsm) {
    S Jrow[n];
    for (int j = 0; j < n; ++j) Jrow[j] = at2(Asm, c, j, tl);
    for (int i = 0; i < n; ++i) {
      S s = at2(Bsm, i, 0, tl) * Jrow[0];
      for (int j = 1; j < n; ++j) s = s + at2(Bsm, i, j, tl) * Jrow[j];
      Pcol[i] = s;
    }
    if (a.noise_mode == NOISE_COVFN) {
      if (a.cov_fn == COV_DIAGONAL) { const S e = eps[c] * a.cov_scale; Pcol[c] = Pcol[c] + e * e; }
      else if (a.cov_fn == COV_OUTER) {
        for (int i = 0; i < n; ++i) Pcol[i] = Pcol[i] + (eps[i] * a.cov_scale) * (eps[c] * a.cov_scale);
      } else Pcol[c] = Pcol[c] + a.cov_scale * a.cov_scale;
    } else if (a.noise_mode == NOISE_EPS_PLUS_Q) {
      if (a.qdiag) Pcol[c] = Pcol[c] + q_own(a);
      else for (int i = 0; i < n; ++i) Pcol[i] = Pcol[i] + a.GQ[i * n + c];
      Pcol[c] = Pcol[c] + eps[c] * eps[c];
    } else if (a.noise_mode == NOISE_Q_ONLY) {
      if (a.qdiag) Pcol[c] = Pcol[c] + q_own(a);
      else for (int i = 0; i < n; ++i) Pcol[i] = Pcol[i] + a.GQ[i * n + c];
    }
    for (int i = 0; i < n; ++i) x[i] = xn[i];
    t = t + a.h;
  }
  // phase 4a: PHt row c (local by symmetry of P) -> smem
  ODEU_HD void phase_pht(const Args& a, S* Bsm) {
    const int L = a.L;
    for (int l = 0; l < L; ++l) {
      S s = S(0.0);
      for (int j = 0; j < n; ++j) s = s + Pcol[j] * a.H[l * n + j];
      ph[l] = s;
      at3(Bsm, 0, c, l, L, tl) = s;
    }
  }
  // phase 4b: S, Cholesky, NLL term, gain row c, residual row c -> smem
  ODEU_HD void phase_gain(const Args& a, const double* y, S* Bsm) {
    const int L = a.L;
    S Sm[n][n], Ls[n][n], inv[n], z[n];
    for (int l = 0; l < L; ++l) {
      S s = S(0.0);
      for (int j = 0; j < n; ++j) s = s + x[j] * a.H[l * n + j];
      dvec[l] = y[l] - s;
      for (int m = 0; m <= l; ++m) {
        S v = S(a.R[l * L + m]);
        for (int i = 0; i < n; ++i) v = v + at3(Bsm, 0, i, m, L, tl) * a.H[l * n + i];
        Sm[l][m] = v;
        Sm[m][l] = v;
      }
    }
    bool all_tiny = true;
    S logdet = S(0.0), quad = S(0.0);
    for (int j = 0; j < L; ++j) {
      S s = Sm[j][j];
      for (int k = 0; k < j; ++k) s = s - Ls[j][k] * Ls[j][k];
      const S dj = d_sqrt(s);
      Ls[j][j] = dj;
      inv[j] = 1.0 / dj;
      all_tiny = all_tiny && (fabs(scalar_ops<S>::val(dj)) < 1e-16);
      logdet = logdet + d_log(d_abs(dj));
      for (int i = j + 1; i < L; ++i) {
        S v = Sm[i][j];
        for (int k = 0; k < j; ++k) v = v - Ls[i][k] * Ls[j][k];
        v = v * inv[j];
        Ls[i][j] = v;
        all_tiny = all_tiny && (fabs(scalar_ops<S>::val(v)) < 1e-16);
      }
    }
    for (int i = 0; i < L; ++i) {
      S s = dvec[i];
      for (int k = 0; k < i; ++k) s = s - Ls[i][k] * z[k];
      z[i] = s * inv[i];
      quad = quad + z[i] * z[i];
    }
    if (c == 0) nll = nll + (quad * 0.5 + logdet + 0.5 * (double)L * 1.8378770664093453);
    S w[n];
    for (int l = 0; l < L; ++l) {
      S s = ph[l];
      for (int k = 0; k < l; ++k) s = s - Ls[l][k] * w[k];
      w[l] = s * inv[l];
    }
    for (int l = L - 1; l >= 0; --l) {
      S s = w[l];
      for (int k = l + 1; k < L; ++k) s = s - Ls[k][l] * Krow[k];
      Krow[l] = all_tiny ? S(0.0) : s * inv[l];
    }
    for (int l = 0; l < L; ++l) {
      S g = ph[l];                       // G = PHt - K S  (rounding-level residual of the Joseph form)
      for (int m = 0; m < L; ++m) g = g - Krow[m] * Sm[m][l];
      at3(Bsm, 1, c, l, L, tl) = Krow[l];
      at3(Bsm, 2, c, l, L, tl) = g;
    }
  }
  // phase 4c: x += K d;  P+[:, c] = (P - K (H P))[:, c] - G K[c, :]^T
  ODEU_HD void phase_update(const Args& a, S* Bsm) {
    const int L = a.L;
    for (int i = 0; i < n; ++i) {
      S s = x[i], p = Pcol[i];
      for (int l = 0; l < L; ++l) {
        const S Kil = at3(Bsm, 1, i, l, L, tl);
        s = s + Kil * dvec[l];
        p = p - Kil * ph[l];
      }
      for (int l = 0; l < L; ++l) p = p - at3(Bsm, 2, i, l, L, tl) * Krow[l];
      x[i] = s;
      Pcol[i] = p;
    }
  }
  ODEU_HD void finish(const Args& a, double* PT) {
    if (!active) return;
    if (c == 0) {
      if (chunk == 0) {
        if (a.nll) a.nll[b] = scalar_ops<S>::val(nll);
        if (a.xT)
          for (int i = 0; i < n; ++i) a.xT[i * a.B + b] = scalar_ops<S>::val(x[i]);
      }
      if constexpr (!std::is_same<S, double>::value) {
        if (a.grad && chunk < a.p_opt) a.grad[(long long)chunk * a.B + b] = nll.d[0];
      }
    }
    if (PT && chunk == 0)
      for (int i = 0; i < n; ++i) PT[((long long)i * n + c) * a.B + b] = scalar_ops<S>::val(Pcol[i]);
  }
};

// dynamic shared memory: 2 * n * n * TB scalars
template <class Ode, class Tab, class S, int TB>
__global__ void __launch_bounds__(Ode::NX * TB)
ekf_coop_kernel(const __grid_constant__ GradArgs<Ode::NX, Ode::NP> a, double* PT) {
  constexpr int n = Ode::NX;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S* Asm = reinterpret_cast<S*>(smem_raw);
  S* Bsm = Asm + (long long)n * n * TB;
  CoopThread<Ode, Tab, S, TB> th;
  const int tl = threadIdx.x % TB, c = threadIdx.x / TB;
  th.init(a, (long long)blockIdx.x * TB + tl, tl, c);
  for (long long step = 0; step < a.T; ++step) {
    th.phase_rk(a, Asm);
    __syncthreads();
    th.phase_m(Asm, Bsm);
    __syncthreads();
    th.phase_p(a, Asm, Bsm);
    __syncthreads();
    if (a.has_obs && a.flags[step]) {
      const long long oi = a.ymap[step];
      double y[n];
      for (int l = 0; l < a.L; ++l)
        y[l] = a.ys_per_traj ? a.ys[(oi * a.L + l) * a.B + th.b] : a.ys[oi * a.L + l];
      th.phase_pht(a, Bsm);
      __syncthreads();
      th.phase_gain(a, y, Bsm);
      __syncthreads();
      th.phase_update(a, Bsm);
      __syncthreads();
    }
  }
  th.finish(a, PT);
}

}  // namespace odeu
