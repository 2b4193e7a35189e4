// Row-parallel EKF kernel (one CTA = a block of trajectories x their rows) for ODE plugins with a sparse right-hand-side Jacobian
// (the Hodgkin-Huxley family; BASELINE config 3: loss and forward-mode gradient).
//
// Same step as everywhere else (predict src/filters/sqrt_ekf.py:92-197 over the RK step
// src/solvers/rksolver.py:113-155, correct :337-376, log-likelihood src/utils.py:109-128), but
// organised so that nothing is evaluated twice:
//
//   * thread (trajectory tl, row r) owns equation r of the system: it evaluates f_r at every RK
//     stage TOGETHER with the non-zero partials d f_r / d x (Ode::row), so each `exp` of the rate
//     functions is computed once per stage (round 1's column-parallel kernel recomputed
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
#include "ekf_grad.cuh"
#include "vec2.cuh"

namespace odeu {

constexpr int ROWS_LMAX = 4;   // largest observation dimension served by this kernel

template <class Ode, class = void> struct has_rows : std::false_type {};
template <class Ode> struct has_rows<Ode, std::void_t<decltype(Ode::ROW_CLASSES)>> : std::true_type {};

// Which stage tangents the propagated solution needs, and where they are kept
// (compile-time functions of the tableau).
template <class Tab>
__host__ __device__ constexpr bool tan_need(int i) {
  bool need[8] = {false, false, false, false, false, false, false, false};
  for (int s = Tab::S - 1; s >= 0; --s) {
    bool nd = Tab::b(1, s) != 0.0;
    for (int k = s + 1; k < Tab::S; ++k)
      if (Tab::a(k, s) != 0.0 && need[k]) nd = true;
    need[s] = nd;
  }
  return i < Tab::S && need[i];
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
// >= 0 shared-memory slot, -1 registers, -2 last needed stage (accumulated straight into J),
// -3 tangent not needed
template <class Tab>
__host__ __device__ constexpr int tan_slot(int i) {
  if (!tan_need<Tab>(i)) return -3;
  if (i == tan_last<Tab>()) return -2;
  int stored = 0;
  for (int k = 0; k <= i; ++k)
    if (tan_need<Tab>(k)) ++stored;
  if (TanSched<Tab>::NREG == 1 && stored == TanSched<Tab>::NSTORED) return -1;
  return stored - 1;
}

// pre-scaled tableau (constant-bank operands of the tangent DFMAs)
template <class Tab, int NXA, int NPA>
void fill_rows_schedule(GradArgs<NXA, NPA>& a) {
  a.rw_last = tan_last<Tab>();
  for (int i = 0; i < 8; ++i) {
    a.rw_need[i] = tan_need<Tab>(i) ? 1 : 0;
    a.rw_slot[i] = tan_slot<Tab>(i);
    a.rw_hb1[i] = i < Tab::S ? a.h * Tab::b(1, i) : 0.0;
    for (int j = 0; j < 8; ++j) a.rw_ha[i][j] = (i < Tab::S && j < Tab::S) ? a.h * Tab::a(i, j) : 0.0;
  }
}

// Shared-memory map (units of S).  DF, P and the stage tangents use a PAIRED layout when S is a
// plain double: elements 2e, 2e+1 of one trajectory are adjacent, so one LDS.128 fetches two
// operands ([e/2][TB][2]); otherwise [e][TB].  X, TH and the measurement exchange are [e][TB].
// smallest stride >= base, multiple of align, with mult * stride = gp (mod bankrow); gp = 0: no padding
constexpr int rows_pad_stride(int base, int mult, int align, int gp, int bankrow) {
  int best = base;
  for (int st = (base + align - 1) / align * align; gp != 0 && st < base + 2 * bankrow; st += align)
    if ((mult * st) % bankrow == gp) { best = st; break; }
  return best;
}

template <class Ode, class Tab, class S, int TB>
struct RowsSmem {
  static constexpr int n = Ode::NX;
  static constexpr int NP2 = n + (n & 1);              // padded leading dimension (even)
  static constexpr int NSLOT = TanSched<Tab>::NSLOT;
  static constexpr int PRS = sizeof(S) == 8 ? 2 : 1;   // elements per 16 bytes (paired layout)
  // Bank placement of the row groups.  The lanes of one trajectory's G row groups are neighbours (see the kernel),
  // so a quarter warp holds 8 / G trajectories (x 16 bytes = 128 / G bytes of banks) times G groups: an access
  // whose address depends on the thread's own row r = g Q + q is conflict free when the group step moves it by
  // GP = 128 / G bytes (mod 128).  The strides below are padded accordingly (Q entries apart = one group apart);
  // without it every own-row / own-column access is a 2-way conflict (ncu: LDS.128 8 wavefronts instead of 4).
  static constexpr int G = Ode::ROW_GROUPS;
  static constexpr int Q = Ode::ROW_CLASSES;
  static constexpr int BANKROW = 128 / (int)sizeof(S);
  static constexpr int GP = (G > 1 && BANKROW % G == 0) ? BANKROW / G : 0;
  static constexpr int pad_stride(int base, int mult, int align) { return rows_pad_stride(base, mult, align, GP, BANKROW); }
  static constexpr int CS = pad_stride(NP2 * TB, Q, PRS);      // column stride of P and of the stage tangents
  static constexpr int XS = pad_stride(TB, Q, 1);              // entry stride of the stage state
  static constexpr int TS = TB;                                // entry stride of the parameters (not padded: the 2-way
                                                               // conflict of the few parameter reads is cheaper than
                                                               // losing the second CTA per SM to 240 doubles)
  static constexpr int RS = pad_stride(ROWS_LMAX * TB, Q, 1);  // row stride of the measurement exchange
  static constexpr int EXN = 3 * n * RS;
  static constexpr int MAT = n * CS;
  static constexpr bool EX_ALIAS = NSLOT >= 2 && MAT >= EXN;
  static constexpr int DFN = ((Ode::NNZ > n ? Ode::NNZ : n) + 1) / 2 * 2;
  static constexpr int o_th = 0;
  static constexpr int o_x = o_th + (Ode::NP + (Ode::NP & 1)) * TS;
  static constexpr int o_df = o_x + NP2 * XS;
  static constexpr int o_p = o_df + DFN * TB + (G - 1) * GP;
  static constexpr int o_ks = o_p + MAT;
  static constexpr int o_ex = EX_ALIAS ? o_ks + MAT : o_ks + NSLOT * MAT;
  static constexpr int total = EX_ALIAS ? o_ks + NSLOT * MAT : o_ex + EXN;
  static constexpr size_t bytes = (size_t)total * sizeof(S);
};

template <class Ode, class Tab, class S, int TB, int LT = 0>
struct RowThread {
  static constexpr int n = Ode::NX;
  static constexpr int NP = Ode::NP;
  static constexpr int Q = Ode::ROW_CLASSES;
  static constexpr int G = Ode::ROW_GROUPS;
  static constexpr int ST = Tab::S;
  static constexpr int LM = ROWS_LMAX;
  static constexpr int PR = sizeof(S) == 8 ? 2 : 1;
  using Args = GradArgs<Ode::NX, Ode::NP>;
  using SM = RowsSmem<Ode, Tab, S, TB>;
  static constexpr int NP2 = SM::NP2;
  static constexpr int CS = SM::CS, XS = SM::XS, TS = SM::TS, RS = SM::RS, GP = SM::GP;
  static_assert(Q * G == n, "rows = groups x classes");

  // ---- per-thread state
  static constexpr int NL = lanes_of<S>::value;   // trajectories carried by one scalar
  int tl, g, q, r;
  int per;         // paired-layout offset of element r
  long long b[NL];
  int chunk;
  bool active[NL];
  double t;
  S x, epsr, nll;
  S fcur;          // k_i of the stage being evaluated (filed into kp[] with a static index)
  S kp[ST];        // own component of the stage derivatives k_j
  S Kreg[n];       // register-resident stage tangent column
  S W[n];          // scratch: Y / row of P
  S ph[LM], Krow[LM], dvec[LM];
  // trajectory-output bookkeeping (NLL-only instantiation, see save_step)
  long long next_save, slot;
  bool obs_fresh;

  // ---- shared memory addressing
  ODEU_HD static constexpr int pe(int e) { return (e / PR) * (TB * PR) + (e % PR); }
  ODEU_HD S* bl(S* sm) const { return sm + tl; }            // plain layout base
  ODEU_HD S* bp(S* sm) const { return sm + tl * PR; }       // paired layout base
  ODEU_HD static void ld2(const S* p, S& u, S& v) {         // elements e (even), e + 1
    if constexpr (PR == 2) {
      const double2 w = *reinterpret_cast<const double2*>(p);
      u = w.x; v = w.y;
    } else {
      u = p[0]; v = p[TB];
    }
  }
  ODEU_HD static void st2(S* p, const S& u, const S& v) {
    if constexpr (PR == 2) {
      *reinterpret_cast<double2*>(p) = make_double2(u, v);
    } else {
      p[0] = u; p[TB] = v;
    }
  }
  ODEU_HD S* Xp(S* sm) const { return bl(sm) + SM::o_x; }
  ODEU_HD S* THp(S* sm) const { return bl(sm) + SM::o_th; }
  ODEU_HD S* DFp(S* sm) const { return bp(sm) + SM::o_df; }
  ODEU_HD S* Prow(S* sm, int j) const { return bp(sm) + SM::o_p + j * CS; }        // + pe(k)
  ODEU_HD S* Kcol(S* sm, int slot, int c) const { return bp(sm) + SM::o_ks + (slot * n + c) * CS; }  // + pe(m)
  ODEU_HD S& EX(S* sm, int which, int i, int l) const {
    return bl(sm)[SM::o_ex + (which * n + i) * RS + l * TB];
  }

  // `unit` = first unit of this lane slot; a scalar with NL lanes also carries unit + TB, ...
  ODEU_HD void init(const Args& a, long long unit, int tl_, int g_, int q_, S* sm) {
    tl = tl_; g = g_; q = q_; r = g * Q + q;
    per = (r / PR) * (TB * PR) + (r % PR);
    const long long total = a.B * (a.p_opt > 0 ? a.p_opt : 1);
    chunk = 0;
    x = S(0.0);
#pragma unroll
    for (int u = 0; u < NL; ++u) {
      const long long un = unit + (long long)u * TB;
      active[u] = un < total;
      b[u] = active[u] ? un % a.B : 0;
      if (u == 0) chunk = active[u] ? (int)(un / a.B) : 0;
      lane_set(x, u, a.x0[r * a.B + b[u]]);
    }
    epsr = S(0.0);
    nll = S(0.0);
    t = a.t0;
    for (int k = r; k < NP; k += n) {
      S v = S(0.0);
#pragma unroll
      for (int u = 0; u < NL; ++u) lane_set(v, u, a.theta ? a.theta[k * a.B + b[u]] : a.theta_shared[k]);
      seed_theta(a, k, v);
      THp(sm)[k * TS] = v;
    }
    seed_x0(a);
    for (int k = 0; k < NP2; ++k) {
      S v = S(k < n ? a.P0s[r * n + k] : 0.0);
      if (a.P0b && k < n) {
#pragma unroll
        for (int u = 0; u < NL; ++u) lane_set(v, u, a.P0b[((long long)r * n + k) * a.B + b[u]]);
      }
      Prow(sm, r)[pe(k)] = v;
    }
    next_save = a.save_interval; slot = 1; obs_fresh = false;
    for (int j = 0; j < ST; ++j) kp[j] = S(0.0);
    for (int m = 0; m < n; ++m) Kreg[m] = S(0.0);
  }
  ODEU_HD void seed_theta(const Args& a, int k, S& v) {
    if constexpr (is_gdual<S>::value) {
      if (chunk < a.p_opt && a.idx[chunk] == k) v.d[0] = 1.0;
    }
  }
  ODEU_HD void seed_x0(const Args& a) {
    if constexpr (is_gdual<S>::value) {
      if (chunk < a.p_opt && a.x0_tan) x.d[0] = a.x0_tan[((long long)chunk * n + r) * a.B + b[0]];
    }
  }

  // ---- stage I, interval A: own component of the stage state  x_i = x + h (ks @ A[i])
  template <int I>
  ODEU_HD void stage_a(const Args& a, S* sm) {
    if constexpr (I < ST) {
      S xi = x;
      if constexpr (I > 0) {
        S s = S(0.0);
        bool first = true;
#pragma unroll
        for (int j = 0; j < I; ++j) {
          if (Tab::a(I, j) != 0.0) {
            s = first ? kp[j] * a.rt_A[I][j] : s + kp[j] * a.rt_A[I][j];
            first = false;
          }
        }
        if (!first) xi = x + s * a.h;
      }
      Xp(sm)[r * XS] = xi;
    }
  }
  ODEU_HD void stage_a_rt(const Args& a, int i, S* sm) {
    switch (i) {
      case 0: stage_a<0>(a, sm); break;
      case 1: stage_a<1>(a, sm); break;
      case 2: stage_a<2>(a, sm); break;
      case 3: stage_a<3>(a, sm); break;
      case 4: stage_a<4>(a, sm); break;
      case 5: stage_a<5>(a, sm); break;
      case 6: stage_a<6>(a, sm); break;
      default: stage_a<7>(a, sm); break;
    }
  }
  // ---- interval B: equation r and its partials at the stage state (one copy of the code)
  ODEU_HD void stage_b(const Args& a, int i, S* sm) {
    S f, df[Q + 2];
    Ode::template row<S>(q, g, t + a.h * a.rt_c[i], Xp(sm), XS, THp(sm), TS, f, df);
    fcur = f;
    if (a.rw_need[i]) {
      S* d = DFp(sm) + Ode::row_off(g, q) * TB + g * GP;     // row offsets are even; group g shifted by g GP
      const int nd = Ode::row_ndep(q);
#pragma unroll
      for (int k = 0; k < Q + 2; ++k)
        if (k < nd) d[pe(k)] = df[k];
    }
  }
  // ---- interval C: tangent column r of stage I;  after the last needed stage: J[:, r]
  template <int SLOT>
  ODEU_HD void add_stage(S* sm, double c) {          // W += c K'_slot[:, r]
    if constexpr (SLOT >= 0) {
      const S* kc = Kcol(sm, SLOT, r);
#pragma unroll
      for (int m = 0; m + 1 < n; m += 2) {
        S u, v;
        ld2(kc + pe(m), u, v);
        W[m] = W[m] + u * c;
        W[m + 1] = W[m + 1] + v * c;
      }
      if constexpr (n & 1) W[n - 1] = W[n - 1] + kc[pe(n - 1)] * c;
    } else {
#pragma unroll
      for (int m = 0; m < n; ++m) W[m] = W[m] + Kreg[m] * c;
    }
  }
  template <int I, int J0, bool BROW>
  ODEU_HD void add_stages(const Args& a, S* sm) {    // j = J0 .. I-1, coefficients h a_Ij or h b1_j
    if constexpr (J0 < I) {
      constexpr double cf = BROW ? Tab::b(1, J0) : Tab::a(I, J0);
      if constexpr (cf != 0.0 && tan_need<Tab>(J0))
        add_stage<tan_slot<Tab>(J0)>(sm, BROW ? a.rw_hb1[J0] : a.rw_ha[I][J0]);
      add_stages<I, J0 + 1, BROW>(a, sm);
    }
  }
  template <int I>
  ODEU_HD void stage_c(const Args& a, S* sm) {
    if constexpr (I < ST) kp[I] = fcur;
    if constexpr (I < ST && tan_need<Tab>(I)) {
      constexpr int SLOT = tan_slot<Tab>(I);
      // Y = e_r + h sum_j a_Ij K'_j[:, r]
#pragma unroll
      for (int m = 0; m < n; ++m) W[m] = S(m == r ? 1.0 : 0.0);
      add_stages<I, 0, false>(a, sm);
      // K'_I[:, r] = Df_I Y   (static sparsity pattern, paired loads of the Jacobian entries)
      S acc[n];
      const S* d = DFp(sm);
#pragma unroll
      for (int gm = 0; gm < G; ++gm)
#pragma unroll
        for (int qm = 0; qm < Q; ++qm) {
          const int off = Ode::row_off(gm, qm);
          const int nd = Ode::row_ndep(qm);
          S s = S(0.0);
#pragma unroll
          for (int k = 0; k + 1 < nd; k += 2) {
            S u, v;
            ld2(d + pe(off + k) + gm * GP, u, v);
            s = (k == 0) ? d_fma(v, W[Ode::row_dep(gm, qm, k + 1)], u * W[Ode::row_dep(gm, qm, k)])
                         : d_fma(v, W[Ode::row_dep(gm, qm, k + 1)], d_fma(u, W[Ode::row_dep(gm, qm, k)], s));
          }
          if (nd & 1) {
            s = (nd == 1) ? d[pe(off + nd - 1) + gm * GP] * W[Ode::row_dep(gm, qm, nd - 1)]
                          : d_fma(d[pe(off + nd - 1) + gm * GP], W[Ode::row_dep(gm, qm, nd - 1)], s);
          }
          acc[gm * Q + qm] = s;
        }
      if constexpr (SLOT >= 0) {
        S* kc = Kcol(sm, SLOT, r);
#pragma unroll
        for (int m = 0; m + 1 < n; m += 2) st2(kc + pe(m), acc[m], acc[m + 1]);
        if constexpr (n & 1) kc[pe(n - 1)] = acc[n - 1];
      } else if constexpr (SLOT == -1) {
#pragma unroll
        for (int m = 0; m < n; ++m) Kreg[m] = acc[m];
      } else {   // last needed stage: J[:, r] = e_r + h sum_j b1_j K'_j[:, r]  -> shared (slot 0)
#pragma unroll
        for (int m = 0; m < n; ++m) W[m] = S(m == r ? 1.0 : 0.0) + acc[m] * a.rw_hb1[I];
        add_stages<I, 0, true>(a, sm);
        S* kc = Kcol(sm, 0, r);
#pragma unroll
        for (int m = 0; m + 1 < n; m += 2) st2(kc + pe(m), W[m], W[m + 1]);
        if constexpr (n & 1) kc[pe(n - 1)] = W[n - 1];
      }
    }
  }
  ODEU_HD void stage_c_rt(const Args& a, int i, S* sm) {
    switch (i) {
      case 0: stage_c<0>(a, sm); break;
      case 1: stage_c<1>(a, sm); break;
      case 2: stage_c<2>(a, sm); break;
      case 3: stage_c<3>(a, sm); break;
      case 4: stage_c<4>(a, sm); break;
      case 5: stage_c<5>(a, sm); break;
      case 6: stage_c<6>(a, sm); break;
      default: stage_c<7>(a, sm); break;
    }
  }
  // ---- after the stages: propagated state (row b[1]), embedded error, time
  ODEU_HD void phase_x(const Args& a, S* sm) {
    S s1 = S(0.0), s0 = S(0.0);
#pragma unroll
    for (int j = 0; j < ST; ++j) {
      if (Tab::b(1, j) != 0.0) s1 = s1 + kp[j] * a.rt_b[1][j];
      if (Tab::b(0, j) != 0.0) s0 = s0 + kp[j] * a.rt_b[0][j];
    }
    const S x1 = x + s1 * a.h;
    const S x0 = x + s0 * a.h;
    epsr = d_abs(x0 - x1);
    x = x1;
    t = t + a.h;
    Xp(sm)[r * XS] = x;
  }
  // ---- row r of M = J P, then of P+ = M J^T + Q        (J(k, j) = Kcol(0, j)[k])
  ODEU_HD void phase_mp(const Args& a, S* sm) {
    S Jr[n];
#pragma unroll
    for (int j = 0; j < n; ++j) Jr[j] = Kcol(sm, 0, j)[per];
    S M[n];
#pragma unroll
    for (int j = 0; j < n; ++j) {
      const S* pj = Prow(sm, j);
#pragma unroll
      for (int k = 0; k + 1 < n; k += 2) {
        S u, v;
        ld2(pj + pe(k), u, v);
        M[k] = (j == 0) ? Jr[j] * u : d_fma(Jr[j], u, M[k]);
        M[k + 1] = (j == 0) ? Jr[j] * v : d_fma(Jr[j], v, M[k + 1]);
      }
      if constexpr (n & 1) {
        const S u = pj[pe(n - 1)];
        M[n - 1] = (j == 0) ? Jr[j] * u : d_fma(Jr[j], u, M[n - 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < n; ++j) {
      const S* kj = Kcol(sm, 0, j);
#pragma unroll
      for (int k = 0; k + 1 < n; k += 2) {
        S u, v;
        ld2(kj + pe(k), u, v);
        W[k] = (j == 0) ? M[j] * u : d_fma(M[j], u, W[k]);
        W[k + 1] = (j == 0) ? M[j] * v : d_fma(M[j], v, W[k + 1]);
      }
      if constexpr (n & 1) {
        const S u = kj[pe(n - 1)];
        W[n - 1] = (j == 0) ? M[j] * u : d_fma(M[j], u, W[n - 1]);
      }
    }
    // process noise (src/filters/sqrt_ekf.py:96-136), row r
    if (a.noise_mode == NOISE_COVFN) {
      if (a.cov_fn == COV_DIAGONAL) { const S e = epsr * a.cov_scale; add_diag(e * e); }
      else if (a.cov_fn == COV_OUTER) DFp(sm)[per] = epsr * a.cov_scale;   // exchanged, added in phase_noise_outer
      else add_diag(S(a.cov_scale * a.cov_scale));
    } else if (a.noise_mode == NOISE_EPS_PLUS_Q) {
      if (a.qdiag) add_diag(q_own(a));
      else {
#pragma unroll
        for (int k = 0; k < n; ++k) W[k] = W[k] + a.GQ[r * n + k];
      }
      add_diag(epsr * epsr);
    } else if (a.noise_mode == NOISE_Q_ONLY) {
      if (a.qdiag) add_diag(q_own(a));
      else {
#pragma unroll
        for (int k = 0; k < n; ++k) W[k] = W[k] + a.GQ[r * n + k];
      }
    }
  }
  // (gamma_sqrt w_r)^2 of the per-trajectory diagonal Q (parameter_sensitivity,
  // run_parameter_estimation.py:750-769); re-read every step (an L1/L2 hit) so that it costs no
  // registers on the runs that do not use it
  ODEU_HD S q_own(const Args& a) const {
    S wv = S(0.0);
#pragma unroll
    for (int u = 0; u < NL; ++u) lane_set(wv, u, a.qdiag[r * a.B + b[u]]);
    if constexpr (is_gdual<S>::value) {
      if (chunk < a.p_opt && a.qdiag_tan) wv.d[0] = a.qdiag_tan[((long long)chunk * n + r) * a.B + b[0]];
    }
    wv = wv * a.q_gamma;
    return wv * wv;
  }
  ODEU_HD void add_diag(const S& v) {
    // unconditional stores: a guarded `if (k == r)` store is turned into a dynamically indexed
    // access by the compiler and drags the whole row into local memory
#pragma unroll
    for (int k = 0; k < n; ++k) {
      const S add = (k == r) ? v : S(0.0);
      W[k] = W[k] + add;
    }
  }
  // rank-one process noise of OuterCovarianceUpdate (needs every component of eps):
  // Q = (s eps)(s eps)^T, NaN when eps == 0 like the reference's 0/0 (outer.py:57-60)
  ODEU_HD void phase_noise_outer(const Args& a, S* sm) {
    if (!(a.noise_mode == NOISE_COVFN && a.cov_fn == COV_OUTER)) return;
    const S* d = DFp(sm);
    const S er = d[per];
    S poison = S(0.0);
#pragma unroll
    for (int u = 0; u < NL; ++u) {
      double ss = 0.0;
#pragma unroll
      for (int k = 0; k < n; ++k) { const double e = lane_get(d[pe(k)], u); ss += e * e; }
      lane_set(poison, u, (ss == 0.0) ? (ss / ss) : 0.0);
    }
#pragma unroll
    for (int k = 0; k < n; ++k) W[k] = W[k] + er * d[pe(k)] + poison;
  }
  // ---- measurement update, row-wise.  PHt row r is local (row r of the symmetric P).
  // LT > 0: observation dimension known at compile time (every predicate below folds away)
  ODEU_HD static int obs_dim(const Args& a) { return LT > 0 ? LT : a.L; }
  ODEU_HD void phase_pht(const Args& a, S* sm) {
    const int L = obs_dim(a);
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
  ODEU_HD void phase_gain(const Args& a, const S* y, S* sm) {
    const int L = obs_dim(a);
    S Sm[LM][LM], Ls[LM][LM], inv[LM], z[LM];
    const S* xs = Xp(sm);
#pragma unroll
    for (int l = 0; l < LM; ++l) {
      if (l < L) {
        if (a.h_sel_all) {     // H row l = unit vector: y_hat_l = x[h_l], S_lm = (P H^T)[h_l][m] + R_lm
          const int hl = a.h_sel[l];
          dvec[l] = y[l] - xs[hl * XS];
#pragma unroll
          for (int m = 0; m < LM; ++m) {
            if (m <= l) {
              const S v = EX(sm, 0, hl, m) + a.R[l * L + m];
              Sm[l][m] = v;
              Sm[m][l] = v;
            }
          }
        } else {
          S s = xs[0] * a.H[l * n];
#pragma unroll
          for (int j = 1; j < n; ++j) s = s + xs[j * XS] * a.H[l * n + j];
          dvec[l] = y[l] - s;
#pragma unroll
          for (int m = 0; m < LM; ++m) {
            if (m <= l) {
              S v = S(a.R[l * L + m]);
#pragma unroll
              for (int i = 0; i < n; ++i) v = v + EX(sm, 0, i, m) * a.H[l * n + i];
              Sm[l][m] = v;
              Sm[m][l] = v;
            }
          }
        }
      }
    }
    publish_obs(a, xs, Sm, L);
    Mask2 all_tiny = mask_true();
#pragma unroll
    for (int j = 0; j < LM; ++j) {
      if (j < L) {
        S s = Sm[j][j];
#pragma unroll
        for (int k = 0; k < LM; ++k) if (k < j) s = s - Ls[j][k] * Ls[j][k];
        inv[j] = d_rsqrt(s);                 // one MUFU + 5 DFMA instead of sqrt + division (~55 instructions,
        const Mask2 z0 = mask_zero(s);       // evaluated redundantly by every row thread).  An exactly singular S
        const S dj = zero_where(z0, s * inv[j]);   // (pivot == 0) gives a zero pivot and a zero column, not
        Ls[j][j] = dj;                             // 0 * inf: the guard below then trips like sqrt_ekf.py:351-353
        mask_and_tiny(all_tiny, dj);
#pragma unroll
        for (int i = 0; i < LM; ++i) {
          if (i > j && i < L) {
            S v = Sm[i][j];
#pragma unroll
            for (int k = 0; k < LM; ++k) if (k < j) v = v - Ls[i][k] * Ls[j][k];
            v = zero_where(z0, v * inv[j]);
            Ls[i][j] = v;
            mask_and_tiny(all_tiny, v);
          }
        }
      }
    }
    if (r == 0) {   // the log-likelihood term is kept by ONE thread of the trajectory (its logs are not free)
      S logdet = S(0.0), quad = S(0.0);
#pragma unroll
      for (int i = 0; i < LM; ++i) {
        if (i < L) {
          S s = dvec[i];
#pragma unroll
          for (int k = 0; k < LM; ++k) if (k < i) s = s - Ls[i][k] * z[k];
          z[i] = s * inv[i];
          quad = quad + z[i] * z[i];
          logdet = logdet + d_log(d_abs(Ls[i][i]));
        }
      }
      nll = nll + (quad * 0.5 + logdet + 0.5 * (double)L * 1.8378770664093453);
    }
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
        Krow[l] = zero_where(all_tiny, s * inv[l]);
      }
    }
#pragma unroll
    for (int l = 0; l < LM; ++l) {
      if (l < L) {
        S gg = ph[l];                     // G = PHt - K S  (rounding-level residual of the Joseph form)
#pragma unroll
        for (int m = 0; m < LM; ++m) if (m < L) gg = gg - Krow[m] * Sm[m][l];
        EX(sm, 1, r, l) = Krow[l];
        ph[l] = gg;                       // from here on ph holds G_r
        EX(sm, 2, r, l) = gg;
      }
    }
  }
  // The PRE-update y_hat = y - d and S that the reference keeps in its state (sqrt_ekf.py:370-372): thread
  // r = 0 of the trajectory writes them into the upcoming trajectory slot and the final-state arrays.
  ODEU_HD void publish_obs(const Args& a, const S* xs, const S (*Sm)[LM], int L) {
    if constexpr (!is_gdual<S>::value && NL == 1) {
      obs_fresh = true;
      if (r != 0 || !active[0]) return;
      if (!(a.save_interval > 0 || a.yhatT || a.ST)) return;
      const bool to_slot = a.save_interval > 0 && next_save <= a.T;
      for (int l = 0; l < L; ++l) {
        double yh;
        if (a.h_sel_all) yh = lane_get(xs[a.h_sel[l] * XS], 0);
        else {
          yh = lane_get(xs[0], 0) * a.H[l * n];
          for (int j = 1; j < n; ++j) yh = yh + lane_get(xs[j * XS], 0) * a.H[l * n + j];
        }
        if (to_slot && a.out_yhat) a.out_yhat[(slot * L + l) * a.B + b[0]] = yh;
        if (a.yhatT) a.yhatT[l * a.B + b[0]] = yh;
        for (int m = 0; m < L; ++m) {
          const double sv = lane_get(Sm[l][m], 0);
          if (to_slot && a.out_S) a.out_S[(slot * L * L + l * L + m) * a.B + b[0]] = sv;
          if (a.ST) a.ST[(l * L + m) * a.B + b[0]] = sv;
        }
      }
    }
  }
  // slot 0 (initial state) and, after a step, the strided save of run_filter.py:219-222: thread (trajectory,
  // row r) writes x_r, eps_r and row r of P (still in W after phase_mp / phase_update)
  ODEU_HD void save_initial(const Args& a, S* sm) {
    if constexpr (!is_gdual<S>::value && NL == 1) {
      if (a.save_interval <= 0 || !active[0]) return;
      const long long B = a.B, bb = b[0];
      if (a.out_x) a.out_x[(long long)r * B + bb] = lane_get(x, 0);
      if (a.out_eps) a.out_eps[(long long)r * B + bb] = 0.0;
      if (a.out_P)
        for (int k = 0; k < n; ++k) a.out_P[((long long)r * n + k) * B + bb] = lane_get(Prow(sm, r)[pe(k)], 0);
      if (r == 0) {
        const int L = a.L;
        for (int l = 0; l < L; ++l) if (a.out_yhat) a.out_yhat[l * B + bb] = 0.0;
        for (int l = 0; l < L * L; ++l) if (a.out_S) a.out_S[l * B + bb] = 0.0;
        if (bb == 0 && a.out_t) a.out_t[0] = t;
      }
    }
  }
  ODEU_HD void save_step(const Args& a, long long step) {
    if constexpr (!is_gdual<S>::value && NL == 1) {
      if (a.save_interval <= 0 || step + 1 != next_save) return;
      if (active[0]) {
        const long long B = a.B, bb = b[0];
        if (a.out_x) a.out_x[(slot * n + r) * B + bb] = lane_get(x, 0);
        if (a.out_eps) a.out_eps[(slot * n + r) * B + bb] = lane_get(epsr, 0);
        if (a.out_P) {
#pragma unroll
          for (int k = 0; k < n; ++k) a.out_P[(slot * n * n + (long long)r * n + k) * B + bb] = lane_get(W[k], 0);
        }
        if (r == 0) {
          const int L = a.L;
          if (!obs_fresh) {      // no update since the previous slot: the state still holds the older values
            for (int l = 0; l < L; ++l)
              if (a.out_yhat) a.out_yhat[(slot * L + l) * B + bb] = a.out_yhat[((slot - 1) * L + l) * B + bb];
            for (int l = 0; l < L * L; ++l)
              if (a.out_S) a.out_S[(slot * L * L + l) * B + bb] = a.out_S[((slot - 1) * L * L + l) * B + bb];
          }
          if (bb == 0 && a.out_t) a.out_t[slot] = t;
        }
      }
      obs_fresh = false;
      ++slot;
      next_save += a.save_interval;
    }
  }
  // x_r += K_r d;  P+[r, :] = P[r, :] - K_r (H P)[:, :] - G_r K^T
  ODEU_HD void phase_update(const Args& a, S* sm) {
    const int L = obs_dim(a);
#pragma unroll
    for (int l = 0; l < LM; ++l) {
      if (l < L) {
        x = d_fma(Krow[l], dvec[l], x);
        const S nk = -Krow[l], ng = -ph[l];
#pragma unroll
        for (int k = 0; k < n; ++k) W[k] = d_fma(ng, EX(sm, 1, k, l), d_fma(nk, EX(sm, 0, k, l), W[k]));
      }
    }
  }
  ODEU_HD void phase_store(S* sm) {
    S* pr = Prow(sm, r);
#pragma unroll
    for (int k = 0; k + 1 < n; k += 2) st2(pr + pe(k), W[k], W[k + 1]);
    if constexpr (n & 1) pr[pe(n - 1)] = W[n - 1];
  }
  ODEU_HD void finish(const Args& a, double* PT, S* sm) {
#pragma unroll
    for (int u = 0; u < NL; ++u) {
      if (!active[u]) continue;
      if (chunk == 0) {
        if (r == 0 && a.nll) a.nll[b[u]] = lane_get(nll, u);
        if (a.xT) a.xT[r * a.B + b[u]] = lane_get(x, u);
        if (a.epsT) a.epsT[r * a.B + b[u]] = lane_get(epsr, u);
        if (a.tT && r == 0 && b[u] == 0) a.tT[0] = t;
        if (PT)
          for (int k = 0; k < n; ++k) PT[((long long)r * n + k) * a.B + b[u]] = lane_get(Prow(sm, r)[pe(k)], u);
      }
    }
    if constexpr (is_gdual<S>::value) {
      if (active[0] && r == 0 && a.grad && chunk < a.p_opt) a.grad[(long long)chunk * a.B + b[0]] = nll.d[0];
    }
  }
};

// One CTA = TB trajectories (or (trajectory, direction) units) x n rows; warp w = row class w,
// lane = (group, trajectory).  Dynamic shared memory: RowsSmem::bytes.
template <class Ode, class Tab, class S, int TB, int MINB, int LT>
__global__ void __launch_bounds__(32 * Ode::ROW_CLASSES, MINB)
ekf_rows_kernel(const __grid_constant__ GradArgs<Ode::NX, Ode::NP> a, double* PT) {
  static_assert(TB * Ode::ROW_GROUPS == 32, "a warp holds every row group of TB trajectories");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S* sm = reinterpret_cast<S*>(smem_raw);
  RowThread<Ode, Tab, S, TB, LT> th;
  const int lane = threadIdx.x & 31;
  constexpr int NL = lanes_of<S>::value;
  // lanes G t .. G t + G - 1 = the row groups of trajectory t: the lanes that read the SAME Jacobian / covariance
  // entries sit next to each other, where the shared-memory pipe merges them into one wavefront (+6 % at C3
  // against the group-major order lane = g * TB + t)
  constexpr int GR = Ode::ROW_GROUPS;
  th.init(a, (long long)blockIdx.x * (TB * NL) + lane / GR, lane / GR, lane % GR, threadIdx.x >> 5, sm);
  __syncthreads();
  th.save_initial(a, sm);
  for (long long step = 0; step < a.T; ++step) {
#pragma unroll 1
    for (int i = 0; i < Tab::S; ++i) {
      th.stage_a_rt(a, i, sm);
      __syncthreads();
      th.stage_b(a, i, sm);
      __syncthreads();
      th.stage_c_rt(a, i, sm);
    }
    th.phase_x(a, sm);
    __syncthreads();
    th.phase_mp(a, sm);
    __syncthreads();
    th.phase_noise_outer(a, sm);
    if (a.has_obs && a.flags[step]) {
      const long long oi = a.ymap[step];
      const int L = LT > 0 ? LT : a.L;
      S y[ROWS_LMAX];
#pragma unroll
      for (int l = 0; l < ROWS_LMAX; ++l) {
        if (l < L) {
          y[l] = S(0.0);
#pragma unroll
          for (int u = 0; u < NL; ++u)
            lane_set(y[l], u, a.ys_per_traj ? a.ys[(oi * L + l) * a.B + th.b[u]] : a.ys[oi * L + l]);
        }
      }
      th.phase_pht(a, sm);
      __syncthreads();
      th.phase_gain(a, y, sm);
      __syncthreads();
      th.phase_update(a, sm);
    }
    th.phase_store(sm);
    th.save_step(a, step);
  }
  __syncthreads();
  th.finish(a, PT, sm);
}

}  // namespace odeu
