// Row-parallel cooperative EKF kernel for ODE plugins with a sparse right-hand-side Jacobian
// (the Hodgkin-Huxley family; BASELINE config 3: loss and forward-mode gradient).
//
// Same step as everywhere else (predict src/filters/sqrt_ekf.py:92-197 over the RK step
// src/solvers/rksolver.py:113-155, correct :337-376, log-likelihood src/utils.py:109-128), but
// organised so that nothing is evaluated twice:
//
//   * thread (trajectory tl, row r) owns equation r of the system: it evaluates f_r at every RK
//     stage TOGETHER with the non-zero partials d f_r / d x (Ode::row), so each `exp` of the rate
//     functions is computed once per stage (the column-parallel kernel, ekf_coop.cuh, recomputed
//     the whole right-hand side in all n column threads);
//   * the same thread owns tangent COLUMN r of the step Jacobian: the stage tangents follow from
//     K'_i[:, r] = Df_i * (e_r + h sum_j a_ij K'_j[:, r]) with the SPARSE Df_i read from shared
//     memory (NNZ entries, 40 instead of 196 at n = 14), J[:, r] = e_r + h sum_j b1_j K'_j[:, r].
//     Tangent stages nobody needs are skipped (RKF45's sixth stage only feeds the error estimate);
//   * P <- J P J^T + Q and the measurement update are done row-wise (thread r owns row r of P,
//     which lives in shared memory between steps).
//
// A warp holds ONE row class q (same code path for all lanes: no divergence in the rate
// functions) of TB trajectories x G = 32 / TB row groups (compartments); lanes of different
// groups read the same Jacobian / covariance entries, which shared memory serves as a broadcast.
// All shared arrays are laid out [entry][TB], trajectory fastest (conflict free).
//
// The scalar type S is `double` (NLL only) or GDual<double, 1> (NLL + one parameter direction):
// the same statements produce value and derivative; the mixed second-order terms d Df / d theta
// come from the one-lane dual over S inside Ode::row.
//
// The per-thread body is a struct with one method per barrier interval, so the identical source
// is replayed sequentially on the host by the test-only emulation (tests/host_emu.cu).
#pragma once
#include "ekf_coop.cuh"

namespace odeu {

constexpr int ROWS_LMAX = 4;   // largest observation dimension served by this kernel

template <class Ode, class = void> struct has_rows : std::false_type {};
template <class Ode> struct has_rows<Ode, std::void_t<decltype(Ode::ROW_CLASSES)>> : std::true_type {};

// Which stage tangents the propagated solution needs, and where they are kept.
template <class Tab>
__host__ __device__ constexpr bool tan_need(int i) {
  bool need[8] = {false, false, false, false, false, false, false, false};
  for (int s = Tab::S - 1; s >= 0; --s) {
    bool nd = Tab::b(1, s) != 0.0;
    for (int k = s + 1; k < Tab::S; ++k)
      if (Tab::a(k, s) != 0.0 && need[k]) nd = true;
    need[s] = nd;
  }
  return need[i];
}
template <class Tab>
__host__ __device__ constexpr int tan_last() {
  int l = 0;
  for (int i = 0; i < Tab::S; ++i)
    if (tan_need<Tab>(i)) l = i;
  return l;
}
template <class Tab>
__host__ __device__ constexpr int tan_nstored() {   // needed stages before the last one
  int c = 0;
  for (int i = 0; i < tan_last<Tab>(); ++i)
    if (tan_need<Tab>(i)) ++c;
  return c;
}
template <class Tab>
struct TanSched {
  static constexpr int NSTORED = tan_nstored<Tab>();
  static constexpr int LAST = tan_last<Tab>();
  static constexpr int NREG = NSTORED >= 2 ? 1 : 0;          // newest stored stage in registers
  static constexpr int NSLOT = (NSTORED - NREG) >= 1 ? (NSTORED - NREG) : 1;   // in shared memory
};

template <class Tab, int NXA, int NPA>
void fill_rows_schedule(GradArgs<NXA, NPA>& a) {
  using TS = TanSched<Tab>;
  int slot = 0, stored = 0;
  a.rw_last = TS::LAST;
  for (int i = 0; i < 8; ++i) {
    a.rw_need[i] = (i < Tab::S && tan_need<Tab>(i)) ? 1 : 0;
    a.rw_slot[i] = -3;
    if (!a.rw_need[i]) continue;
    if (i == TS::LAST) { a.rw_slot[i] = -2; continue; }        // accumulated straight into J
    ++stored;
    if (TS::NREG == 1 && stored == TS::NSTORED) a.rw_slot[i] = -1;   // register stage
    else a.rw_slot[i] = slot++;
  }
}

template <class Ode, class Tab, class S, int TB>
struct RowsSmem {
  static constexpr int n = Ode::NX;
  static constexpr int NSLOT = TanSched<Tab>::NSLOT;
  static constexpr int EXN = 3 * n * ROWS_LMAX;
  static constexpr bool EX_ALIAS = NSLOT >= 2 && n * n >= EXN;
  static constexpr long long o_th = 0;
  static constexpr long long o_x = o_th + (long long)Ode::NP * TB;
  static constexpr long long o_df = o_x + (long long)n * TB;
  static constexpr long long o_p = o_df + (long long)(Ode::NNZ > n ? Ode::NNZ : n) * TB;
  static constexpr long long o_ks = o_p + (long long)n * n * TB;
  static constexpr long long o_ex = EX_ALIAS ? o_ks + (long long)n * n * TB : o_ks + (long long)NSLOT * n * n * TB;
  static constexpr long long total = EX_ALIAS ? o_ks + (long long)NSLOT * n * n * TB : o_ex + (long long)EXN * TB;
  static constexpr size_t bytes = (size_t)total * sizeof(S);
};

template <class Ode, class Tab, class S, int TB>
struct RowThread {
  static constexpr int n = Ode::NX;
  static constexpr int NP = Ode::NP;
  static constexpr int Q = Ode::ROW_CLASSES;
  static constexpr int G = Ode::ROW_GROUPS;
  static constexpr int ST = 8;
  static constexpr int LM = ROWS_LMAX;
  using Args = GradArgs<Ode::NX, Ode::NP>;
  using SM = RowsSmem<Ode, Tab, S, TB>;
  static_assert(Q * G == n, "rows = groups x classes");

  // ---- per-thread state
  int tl, g, q, r;
  long long b;
  int chunk;
  bool active;
  double t;
  S x, epsr, nll;
  S kp[ST];        // own component of the stage derivatives k_j
  S Kreg[n];       // register-resident stage tangent column
  S W[n];          // scratch: Y / tangent product / row of M / row of P
  S ph[LM], Krow[LM], dvec[LM];

  // ---- shared memory views (trajectory index fastest)
  ODEU_HD S* TH(S* sm) const { return sm + SM::o_th + tl; }
  ODEU_HD S* X(S* sm) const { return sm + SM::o_x + tl; }
  ODEU_HD S* DF(S* sm) const { return sm + SM::o_df + tl; }
  ODEU_HD S& P(S* sm, int i, int j) const { return sm[SM::o_p + ((long long)i * n + j) * TB + tl]; }
  ODEU_HD S& KS(S* sm, int slot, int i, int j) const {
    return sm[SM::o_ks + (((long long)slot * n + i) * n + j) * TB + tl];
  }
  ODEU_HD S& EX(S* sm, int which, int i, int l) const {
    return sm[SM::o_ex + (((long long)which * n + i) * LM + l) * TB + tl];
  }

  ODEU_HD void init(const Args& a, long long unit, int tl_, int g_, int q_, S* sm) {
    tl = tl_; g = g_; q = q_; r = g * Q + q;
    const long long total = a.B * (a.p_opt > 0 ? a.p_opt : 1);
    active = unit < total;
    b = active ? unit % a.B : 0;
    chunk = active ? (int)(unit / a.B) : 0;
    x = S(a.x0[r * a.B + b]);
    epsr = S(0.0);
    nll = S(0.0);
    t = a.t0;
    for (int k = r; k < NP; k += n) {
      S v = S(a.theta ? a.theta[k * a.B + b] : a.theta_shared[k]);
      seed_theta(a, k, v);
      TH(sm)[k * TB] = v;
    }
    seed_x0(a);
    for (int k = 0; k < n; ++k) P(sm, r, k) = S(a.P0s[r * n + k]);
    for (int j = 0; j < ST; ++j) kp[j] = S(0.0);
    for (int m = 0; m < n; ++m) Kreg[m] = S(0.0);
  }
  ODEU_HD void seed_theta(const Args& a, int k, S& v) {
    if constexpr (!std::is_same<S, double>::value) {
      if (chunk < a.p_opt && a.idx[chunk] == k) v.d[0] = 1.0;
    }
  }
  ODEU_HD void seed_x0(const Args& a) {
    if constexpr (!std::is_same<S, double>::value) {
      if (chunk < a.p_opt && a.x0_tan) x.d[0] = a.x0_tan[((long long)chunk * n + r) * a.B + b];
    }
  }

  // ---- stage i, interval A: own component of the stage state  x_i = x + h (ks @ A[i])
  ODEU_HD void stage_a(const Args& a, int i, S* sm) {
    S xi = x;
    if (i > 0) {
      S s = S(0.0);
#pragma unroll
      for (int j = 0; j < ST; ++j)
        if (j < i && a.rt_A[i][j] != 0.0) s = s + kp[j] * a.rt_A[i][j];
      xi = x + s * a.h;
    }
    X(sm)[r * TB] = xi;
  }
  // ---- interval B: equation r and its partials at the stage state
  ODEU_HD void stage_b(const Args& a, int i, S* sm) {
    S f, df[Q + 2];
    Ode::template row<S>(q, g, t + a.h * a.rt_c[i], X(sm), TB, TH(sm), TB, f, df);
#pragma unroll
    for (int j = 0; j < ST; ++j)
      if (j == i) kp[j] = f;
    if (a.rw_need[i]) {
      S* d = DF(sm) + (long long)Ode::row_off(g, q) * TB;
      const int nd = Ode::row_ndep(q);
#pragma unroll
      for (int k = 0; k < Q + 2; ++k)
        if (k < nd) d[k * TB] = df[k];
    }
  }
  // ---- interval C: tangent column r of stage i;  after the last needed stage: J[:, r]
  ODEU_HD void stage_c(const Args& a, int i, S* sm) {
    if (!a.rw_need[i]) return;
    // Y = e_r + h sum_j a_ij K'_j[:, r]
#pragma unroll
    for (int m = 0; m < n; ++m) W[m] = S(m == r ? 1.0 : 0.0);
#pragma unroll 1
    for (int j = 0; j < i; ++j) {
      const double c = a.h * a.rt_A[i][j];
      if (c == 0.0 || !a.rw_need[j]) continue;
      add_stage(a, sm, j, c);
    }
    // K'_i[:, r] = Df_i Y   (static sparsity pattern)
    S acc[n];
    const S* d = DF(sm);
#pragma unroll
    for (int gm = 0; gm < G; ++gm)
#pragma unroll
      for (int qm = 0; qm < Q; ++qm) {
        S s = d[(long long)Ode::row_off(gm, qm) * TB] * W[Ode::row_dep(gm, qm, 0)];
#pragma unroll
        for (int k = 1; k < Ode::row_ndep(qm); ++k)
          s = s + d[(long long)(Ode::row_off(gm, qm) + k) * TB] * W[Ode::row_dep(gm, qm, k)];
        acc[gm * Q + qm] = s;
      }
    const int slot = a.rw_slot[i];
    if (slot >= 0) {
#pragma unroll
      for (int m = 0; m < n; ++m) KS(sm, slot, m, r) = acc[m];
    } else if (slot == -1) {
#pragma unroll
      for (int m = 0; m < n; ++m) Kreg[m] = acc[m];
    } else {   // last needed stage: J[:, r] = e_r + h sum_j b1_j K'_j[:, r]  -> shared (slot 0)
      const double cl = a.h * a.rt_b[1][i];
#pragma unroll
      for (int m = 0; m < n; ++m) W[m] = S(m == r ? 1.0 : 0.0) + acc[m] * cl;
#pragma unroll 1
      for (int j = 0; j < i; ++j) {
        const double c = a.h * a.rt_b[1][j];
        if (c == 0.0 || !a.rw_need[j]) continue;
        add_stage(a, sm, j, c);
      }
#pragma unroll
      for (int m = 0; m < n; ++m) KS(sm, 0, m, r) = W[m];
    }
  }
  ODEU_HD void add_stage(const Args& a, S* sm, int j, double c) {
    const int slot = a.rw_slot[j];
    if (slot >= 0) {
#pragma unroll
      for (int m = 0; m < n; ++m) W[m] = W[m] + KS(sm, slot, m, r) * c;
    } else {
#pragma unroll
      for (int m = 0; m < n; ++m) W[m] = W[m] + Kreg[m] * c;
    }
  }
  // ---- after the stages: propagated state (row b[1]), embedded error, time
  ODEU_HD void phase_x(const Args& a, S* sm) {
    S s1 = S(0.0), s0 = S(0.0);
#pragma unroll
    for (int j = 0; j < ST; ++j) {
      if (j < a.rt_S) {
        if (a.rt_b[1][j] != 0.0) s1 = s1 + kp[j] * a.rt_b[1][j];
        if (a.rt_b[0][j] != 0.0) s0 = s0 + kp[j] * a.rt_b[0][j];
      }
    }
    const S x1 = x + s1 * a.h;
    const S x0 = x + s0 * a.h;
    epsr = d_abs(x0 - x1);
    x = x1;
    t = t + a.h;
    X(sm)[r * TB] = x;
  }
  // ---- row r of M = J P, then of P+ = M J^T + Q
  ODEU_HD void phase_mp(const Args& a, S* sm) {
    S Jr[n];
#pragma unroll
    for (int j = 0; j < n; ++j) Jr[j] = KS(sm, 0, r, j);
    S M[n];
#pragma unroll
    for (int k = 0; k < n; ++k) M[k] = Jr[0] * P(sm, 0, k);
#pragma unroll
    for (int j = 1; j < n; ++j)
#pragma unroll
      for (int k = 0; k < n; ++k) M[k] = M[k] + Jr[j] * P(sm, j, k);
#pragma unroll
    for (int k = 0; k < n; ++k) {
      S s = M[0] * KS(sm, 0, k, 0);
#pragma unroll
      for (int j = 1; j < n; ++j) s = s + M[j] * KS(sm, 0, k, j);
      W[k] = s;
    }
    // process noise (src/filters/sqrt_ekf.py:96-136), row r
    if (a.noise_mode == NOISE_COVFN) {
      if (a.cov_fn == COV_DIAGONAL) { const S e = epsr * a.cov_scale; add_diag(e * e); }
      else if (a.cov_fn == COV_OUTER) DF(sm)[r * TB] = epsr * a.cov_scale;   // exchanged, added in phase_noise_outer
      else add_diag(S(a.cov_scale * a.cov_scale));
    } else if (a.noise_mode == NOISE_EPS_PLUS_Q) {
#pragma unroll
      for (int k = 0; k < n; ++k) W[k] = W[k] + a.GQ[r * n + k];
      add_diag(epsr * epsr);
    } else if (a.noise_mode == NOISE_Q_ONLY) {
#pragma unroll
      for (int k = 0; k < n; ++k) W[k] = W[k] + a.GQ[r * n + k];
    }
  }
  ODEU_HD void add_diag(const S& v) {
#pragma unroll
    for (int k = 0; k < n; ++k)
      if (k == r) W[k] = W[k] + v;
  }
  // rank-one process noise of OuterCovarianceUpdate (needs every component of eps):
  // Q = (s eps)(s eps)^T, NaN when eps == 0 like the reference's 0/0 (outer.py:57-60)
  ODEU_HD void phase_noise_outer(const Args& a, S* sm) {
    if (!(a.noise_mode == NOISE_COVFN && a.cov_fn == COV_OUTER)) return;
    const S er = DF(sm)[r * TB];
    double ss = 0.0;
#pragma unroll
    for (int k = 0; k < n; ++k) { const double e = scalar_ops<S>::val(DF(sm)[k * TB]); ss += e * e; }
    const double poison = (ss == 0.0) ? (ss / ss) : 0.0;
#pragma unroll
    for (int k = 0; k < n; ++k) W[k] = W[k] + er * DF(sm)[k * TB] + poison;
  }
  // ---- measurement update, row-wise.  PHt row r is local (row r of the symmetric P).
  ODEU_HD void phase_pht(const Args& a, S* sm) {
    const int L = a.L;
#pragma unroll
    for (int l = 0; l < LM; ++l) {
      if (l < L) {
        S s = W[0] * a.H[l * n];
#pragma unroll
        for (int j = 1; j < n; ++j) s = s + W[j] * a.H[l * n + j];
        ph[l] = s;
        EX(sm, 0, r, l) = s;
      }
    }
  }
  ODEU_HD void phase_gain(const Args& a, const double* y, S* sm) {
    const int L = a.L;
    S Sm[LM][LM], Ls[LM][LM], inv[LM], z[LM];
#pragma unroll
    for (int l = 0; l < LM; ++l) {
      if (l < L) {
        S s = S(0.0);
        for (int j = 0; j < n; ++j) s = s + X(sm)[j * TB] * a.H[l * n + j];
        dvec[l] = y[l] - s;
#pragma unroll
        for (int m = 0; m < LM; ++m) {
          if (m <= l) {
            S v = S(a.R[l * L + m]);
            for (int i = 0; i < n; ++i) v = v + EX(sm, 0, i, m) * a.H[l * n + i];
            Sm[l][m] = v;
            Sm[m][l] = v;
          }
        }
      }
    }
    bool all_tiny = true;
    S logdet = S(0.0), quad = S(0.0);
#pragma unroll
    for (int j = 0; j < LM; ++j) {
      if (j < L) {
        S s = Sm[j][j];
#pragma unroll
        for (int k = 0; k < LM; ++k) if (k < j) s = s - Ls[j][k] * Ls[j][k];
        const S dj = d_sqrt(s);
        Ls[j][j] = dj;
        inv[j] = 1.0 / dj;
        all_tiny = all_tiny && (fabs(scalar_ops<S>::val(dj)) < 1e-16);
        logdet = logdet + d_log(d_abs(dj));
#pragma unroll
        for (int i = 0; i < LM; ++i) {
          if (i > j && i < L) {
            S v = Sm[i][j];
#pragma unroll
            for (int k = 0; k < LM; ++k) if (k < j) v = v - Ls[i][k] * Ls[j][k];
            v = v * inv[j];
            Ls[i][j] = v;
            all_tiny = all_tiny && (fabs(scalar_ops<S>::val(v)) < 1e-16);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < LM; ++i) {
      if (i < L) {
        S s = dvec[i];
#pragma unroll
        for (int k = 0; k < LM; ++k) if (k < i) s = s - Ls[i][k] * z[k];
        z[i] = s * inv[i];
        quad = quad + z[i] * z[i];
      }
    }
    if (r == 0) nll = nll + (quad * 0.5 + logdet + 0.5 * (double)L * 1.8378770664093453);
    S w[LM];
#pragma unroll
    for (int l = 0; l < LM; ++l) {
      if (l < L) {
        S s = ph[l];
#pragma unroll
        for (int k = 0; k < LM; ++k) if (k < l) s = s - Ls[l][k] * w[k];
        w[l] = s * inv[l];
      }
    }
#pragma unroll
    for (int l = LM - 1; l >= 0; --l) {
      if (l < L) {
        S s = w[l];
#pragma unroll
        for (int k = 0; k < LM; ++k) if (k > l && k < L) s = s - Ls[k][l] * Krow[k];
        Krow[l] = all_tiny ? S(0.0) : s * inv[l];
      }
    }
#pragma unroll
    for (int l = 0; l < LM; ++l) {
      if (l < L) {
        S gg = ph[l];                     // G = PHt - K S  (rounding-level residual of the Joseph form)
#pragma unroll
        for (int m = 0; m < LM; ++m) if (m < L) gg = gg - Krow[m] * Sm[m][l];
        EX(sm, 1, r, l) = Krow[l];
        EX(sm, 2, r, l) = gg;
      }
    }
  }
  // x_r += K_r d;  P+[r, :] = P[r, :] - K_r (H P)[:, :] - G_r K^T
  ODEU_HD void phase_update(const Args& a, S* sm) {
    const int L = a.L;
#pragma unroll
    for (int l = 0; l < LM; ++l) {
      if (l < L) {
        x = x + Krow[l] * dvec[l];
#pragma unroll
        for (int k = 0; k < n; ++k) W[k] = W[k] - Krow[l] * EX(sm, 0, k, l) - ph_g(sm, l) * EX(sm, 1, k, l);
      }
    }
  }
  ODEU_HD S ph_g(S* sm, int l) const { return EX(sm, 2, r, l); }
  ODEU_HD void phase_store(S* sm) {
#pragma unroll
    for (int k = 0; k < n; ++k) P(sm, r, k) = W[k];
  }
  ODEU_HD void finish(const Args& a, double* PT, S* sm) {
    if (!active) return;
    if (chunk == 0) {
      if (r == 0 && a.nll) a.nll[b] = scalar_ops<S>::val(nll);
      if (a.xT) a.xT[r * a.B + b] = scalar_ops<S>::val(x);
      if (PT)
        for (int k = 0; k < n; ++k) PT[((long long)r * n + k) * a.B + b] = scalar_ops<S>::val(P(sm, r, k));
    }
    if constexpr (!std::is_same<S, double>::value) {
      if (r == 0 && a.grad && chunk < a.p_opt) a.grad[(long long)chunk * a.B + b] = nll.d[0];
    }
  }
};

// One CTA = TB trajectories (or (trajectory, direction) units) x n rows; warp w = row class w,
// lane = (group, trajectory).  Dynamic shared memory: RowsSmem::bytes.
template <class Ode, class Tab, class S, int TB, int MINB>
__global__ void __launch_bounds__(32 * Ode::ROW_CLASSES, MINB)
ekf_rows_kernel(const __grid_constant__ GradArgs<Ode::NX, Ode::NP> a, double* PT) {
  static_assert(TB * Ode::ROW_GROUPS == 32, "a warp holds every row group of TB trajectories");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S* sm = reinterpret_cast<S*>(smem_raw);
  RowThread<Ode, Tab, S, TB> th;
  const int lane = threadIdx.x & 31;
  th.init(a, (long long)blockIdx.x * TB + lane % TB, lane % TB, lane / TB, threadIdx.x >> 5, sm);
  __syncthreads();
  const int St = a.rt_S;
  for (long long step = 0; step < a.T; ++step) {
#pragma unroll 1
    for (int i = 0; i < St; ++i) {
      th.stage_a(a, i, sm);
      __syncthreads();
      th.stage_b(a, i, sm);
      __syncthreads();
      th.stage_c(a, i, sm);
    }
    th.phase_x(a, sm);
    __syncthreads();
    th.phase_mp(a, sm);
    __syncthreads();
    th.phase_noise_outer(a, sm);
    if (a.has_obs && a.flags[step]) {
      const long long oi = a.ymap[step];
      double y[ROWS_LMAX];
#pragma unroll
      for (int l = 0; l < ROWS_LMAX; ++l)
        if (l < a.L) y[l] = a.ys_per_traj ? a.ys[(oi * a.L + l) * a.B + th.b] : a.ys[oi * a.L + l];
      th.phase_pht(a, sm);
      __syncthreads();
      th.phase_gain(a, y, sm);
      __syncthreads();
      th.phase_update(a, sm);
    }
    th.phase_store(sm);
  }
  __syncthreads();
  th.finish(a, PT, sm);
}

}  // namespace odeu
