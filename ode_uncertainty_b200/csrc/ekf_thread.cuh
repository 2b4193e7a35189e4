// Whole-trajectory EKF kernel, one thread per trajectory (small states: everything in registers).
//
// Replaces the reference's time loop `lax.scan(scan_wrapper, ...)` (scripts/run_filter.py:204-217)
// and `lax.scan(nll_step, ...)` (scripts/run_parameter_estimation.py:782-794) for a batch of B
// independent trajectories / parameter sets: T steps run inside ONE launch, the state never
// leaves registers, and only the strided saves (save_interval) and the final state touch HBM.
//
// HBM layout (all float64, batch index fastest so every access is coalesced across a warp):
//   x0      [n][B]          P0      [n*n][B] (optional, else shared P0 in the arguments)
//   theta   [NP][B]         ys      [T_obs][L] shared or [T_obs][L][B] per trajectory
//   out_x   [T_save][n][B]  out_eps [T_save][n][B]   out_P [T_save][n*n][B]
//   out_yhat[T_save][L][B]  out_S   [T_save][L*L][B] out_t [T_save]
//   xT [n][B]  epsT [n][B]  PT [n*n][B]  yhatT [L][B]  ST [L*L][B]  nll [B]  tT [1]
#pragma once
#include "ekf_core.cuh"
#include "ekf_sqrt.cuh"
#include "dirk.cuh"

namespace odeu {

template <int NX, int NP>
struct EkfArgs {
  long long B;
  long long T;
  double t0, h;
  int L;
  int noise_mode, cov_fn;
  double cov_scale;
  long long save_interval;  // 0: no trajectory output
  int ys_per_traj;
  int has_obs;
  int skip_predict;       // generic variant only: measurement update without the RK/covariance step
  // device pointers
  const double* x0;
  const double* P0;       // nullable
  const double* theta;    // nullable -> theta_shared
  const double* ys;       // nullable when has_obs == 0
  const unsigned char* flags;
  const long long* ymap;
  const double* scale_b;  // nullable: per-trajectory cov_scale (calibration sweep)
  int nan_to_num;         // per-step NLL terms through nan_to_num
  double* xT; double* epsT; double* PT; double* yhatT; double* ST; double* nll; double* tT;
  double* out_t; double* out_x; double* out_eps; double* out_P; double* out_yhat; double* out_S;
  // small shared matrices, by value (constant bank)
  double P0s[NX * NX];
  double GQ[NX * NX];     // gamma * Q_sqrt Q_sqrt^T
  double H[NX * NX];      // [L][n]
  double R[NX * NX];      // [L][L] = R_sqrt R_sqrt^T
  double theta_shared[NP];
  ScaledTableau st;       // h * a_ij, h * b_1j
  // ---- guard mode "reference" (factor form, ekf_sqrt.cuh; small systems only)
  static constexpr int NF = (NX <= 4) ? NX * NX : 1;
  int guard_verbatim;     // 1: `all(S_sqrt < 1e-16)` as written (sqrt_ekf.py:351); 0: intended |.|
  const double* P0f_b;    // nullable: per-trajectory factor [n*n][B] (resume)
  double* PsT;            // nullable: final factor [n*n][B]
  double* out_Ps;         // nullable: factor per saved slot [T_save][n*n][B]
  long long* guard_counts;  // nullable: [2][B] steps the guard fired / steps the two predicates differ
  double P0f[NF];         // shared P0_sqrt
  double GQs[NF];         // gamma_sqrt * Q_sqrt
  double Rs[NF];          // R_sqrt [L][L]
};

template <int n>
ODEU_HD void save_slot(long long slot, long long B, long long b, int L,
                                          const double* x, const double* eps,
                                          const double (*P)[n], bool obs_fresh, double* out_x,
                                          double* out_eps, double* out_P, double* out_yhat,
                                          double* out_S) {
  constexpr int U = (n <= 4) ? n : 1;
  // one 64-bit multiply per array, then the pointer walks with stride B: the indexed form
  // (slot * n + i) * B + b cost ~9 instructions per store (ncu: +138 instructions per saved step for
  // 15 stores, and the streaming variant's time follows its instruction count)
  if (out_x) {
    double* p = out_x + slot * n * B + b;
#pragma unroll U
    for (int i = 0; i < n; ++i) { *p = x[i]; p += B; }
  }
  if (out_eps) {
    double* p = out_eps + slot * n * B + b;
#pragma unroll U
    for (int i = 0; i < n; ++i) { *p = eps[i]; p += B; }
  }
  if (out_P) {
    double* p = out_P + slot * (n * n) * B + b;
#pragma unroll U
    for (int i = 0; i < n; ++i)
#pragma unroll U
      for (int j = 0; j < n; ++j) { *p = P[i][j]; p += B; }
  }
  // y_hat / S of the most recent measurement update were already written into this slot by
  // ObsSink; when no update happened since the previous slot the reference state still holds
  // the older values (zeros before the first update): carry them forward.
  if (!obs_fresh) {
    for (int l = 0; l < L; ++l)
      if (out_yhat) out_yhat[(slot * L + l) * B + b] = slot > 0 ? out_yhat[((slot - 1) * L + l) * B + b] : 0.0;
    for (int l = 0; l < L * L; ++l)
      if (out_S) out_S[(slot * L * L + l) * B + b] = slot > 0 ? out_S[((slot - 1) * L * L + l) * B + b] : 0.0;
  }
}

template <int n>
ODEU_HD void save_factor_slot(long long slot, long long B, long long b, const double (*Ps)[n], double* out_Ps) {
  if (!out_Ps) return;
  double* p = out_Ps + slot * (n * n) * B + b;
#pragma unroll
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int j = 0; j < n; ++j) { *p = Ps[i][j]; p += B; }
}

// The whole life of trajectory `b`.  __host__ __device__ so the identical source can be
// exercised on the CPU by the test-only host emulation (tests/host_emu.cu); the product only
// ever calls it from the kernel below.
// LK selects the measurement-update code at compile time: -1 = generic (run-time L, dense H),
// 0 = prediction only, 1..n = H = [I_LK 0] (correct_step_lead).
//
// Time segments: the run may be cut into segments [step0, step1) executed by different warps
// (dynamic scheduler below).  Between segments the per-trajectory state (x, P, nll; t per
// 32-trajectory block) lives in the caller-provided workspace `ws`, laid out batch-minor like
// every other array; it is read with L2-only loads because another SM wrote it.
ODEU_HD double ws_load(const double* p) {
#ifdef __CUDA_ARCH__
  return __ldcg(p);
#else
  return *p;
#endif
}

// ---------------------------------------------------------------------------------------------
// Bulk-async staging of the PER-TRAJECTORY observation stream (north_star: "observations ... streamed
// in with ... TMA-staged loads").  The stream is ys [T_obs][L][B]: for one warp (32 consecutive
// trajectories) and one observation row it is a contiguous 256-byte line.  A warp stages OBS_CH steps
// at a time into its own shared-memory ring (2 buffers x OBS_CH x L lines) with `cp.async.bulk`
// (SASS UBLKCP) completing on an mbarrier (SYNCS): lane i of the warp looks up flag / index map of
// step i of the chunk and issues the L line copies of that step, one lane arms the barrier with the
// byte count, and the chunk after the current one is always in flight while the current one is
// consumed.  Device only; needs B % 32 == 0 (whole, 256-byte aligned lines).
constexpr int OBS_CH = 8;
struct ObsStage {
  double* buf;                  // [2][OBS_CH][L][32]
  unsigned long long* bar;      // [2] mbarriers
  unsigned* parity;             // bit k: phase parity to wait for on barrier k (persists across work items)
};
#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
#endif

struct Segment {
  long long step0, step1;   // steps of this segment
  bool first, last;         // first: state comes from x0/P0/t0; last: final outputs are written
  double* ws;               // [n + n*n + 2][B] state (x, P or factor, nll, guard counters), then t per
                            // block of 32 trajectories
  const long long* last_obs;  // nullable: last step with an observation, found once per launch
                            // (find_last_obs_kernel); null = this trajectory scans the flags itself
  const ObsStage* stg;      // nullable: shared-memory staging of per-trajectory observations (device only)
};

// jnp.nan_to_num of one log-likelihood term (calibration sweep, :218 of the calibration script)
ODEU_HD double nll_term(double v, int nan_to_num) {
  if (!nan_to_num) return v;
  if (v != v) return 0.0;
  if (v > 1.7976931348623157e308) return 1.7976931348623157e308;
  if (v < -1.7976931348623157e308) return -1.7976931348623157e308;
  return v;
}

// SQ selects the covariance representation: 0 = full P (intended guard), 1 = factor form with the
// structured QRs (H = [I_LK 0], lower-triangular R_sqrt, diagonal process-noise block), 2 = factor
// form, generic.  In the factor forms `P` below holds P_sqrt.
#ifdef __CUDA_ARCH__
// issue the line copies of the OBS_CH steps starting at `s_begin` into ring buffer `k` (whole warp)
template <int LK, int NXA, int NPA>
__device__ __forceinline__ void stage_issue(const EkfArgs<NXA, NPA>& a, const Segment& sg, long long s_begin, int k, long long b0) {
  const int lane = threadIdx.x & 31;
  __syncwarp();                                          // every lane is done reading this buffer
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  bool fl = false;
  long long oi = 0;
  const long long s = s_begin + lane;
  if (lane < OBS_CH && s < sg.step1) {
    fl = a.flags[s] != 0;
    if (fl) oi = a.ymap[s];
  }
  const unsigned m = __ballot_sync(0xffffffffu, fl);
  if (lane == 0) mbar_expect_tx(sg.stg->bar + k, (unsigned)__popc(m) * LK * 256u);
  __syncwarp();
  if (fl) {
#pragma unroll
    for (int l = 0; l < LK; ++l)
      bulk_g2s(sg.stg->buf + ((k * OBS_CH + lane) * LK + l) * 32, a.ys + (oi * LK + l) * a.B + b0, 256u, sg.stg->bar + k);
  }
}
#endif

template <class Ode, class Tab, int KC, int LK, int SQ = 0, bool STG = false>
ODEU_HD void ekf_trajectory(const EkfArgs<Ode::NX, Ode::NP>& a, const long long b, const Segment& sg) {
  constexpr int n = Ode::NX;
  const double cov_scale = a.scale_b ? a.scale_b[b] : a.cov_scale;
  constexpr int NP = Ode::NP;
  constexpr int U = (n <= 4) ? n : 1;
  const long long B = a.B;
  const int L = a.L;

  double x[n], eps[n], P[n][n], th[NP];
  double t, nll;
  LogProd lp;              // deferred log-determinant terms of this segment
  lp.reset();
  GuardCount gc = {0, 0};
  if (sg.first) {
#pragma unroll U
    for (int i = 0; i < n; ++i) x[i] = a.x0[i * B + b];
    if constexpr (SQ != 0) {
#pragma unroll U
      for (int i = 0; i < n; ++i)
#pragma unroll U
        for (int j = 0; j < n; ++j) P[i][j] = a.P0f_b ? a.P0f_b[(i * n + j) * B + b] : a.P0f[i * n + j];
    } else {
#pragma unroll U
    for (int i = 0; i < n; ++i)
#pragma unroll U
      for (int j = 0; j < n; ++j) P[i][j] = a.P0 ? a.P0[(i * n + j) * B + b] : a.P0s[i * n + j];
    }
    // final-state y_hat / S start at zero like SQRT_EKF.init_state (sqrt_ekf.py:80-82)
    for (int l = 0; l < L; ++l)
      if (a.yhatT) a.yhatT[l * B + b] = 0.0;
    for (int l = 0; l < L * L; ++l)
      if (a.ST) a.ST[l * B + b] = 0.0;
    t = a.t0;
    nll = 0.0;
  } else {
#pragma unroll U
    for (int i = 0; i < n; ++i) x[i] = ws_load(sg.ws + i * B + b);
    if constexpr (SQ != 0) {
#pragma unroll U
      for (int i = 0; i < n; ++i)
#pragma unroll U
        for (int j = 0; j < n; ++j) P[i][j] = ws_load(sg.ws + (n + i * n + j) * B + b);
      const long long packed = LogProd::to_bits(ws_load(sg.ws + (n + n * n + 1) * B + b));
      gc.fired = (int)(packed & 0xffffffffLL);
      gc.mismatch = (int)(packed >> 32);
    } else {
#pragma unroll U
    for (int i = 0; i < n; ++i)
#pragma unroll U
      for (int j = 0; j <= i; ++j) {
        const double v = ws_load(sg.ws + (n + i * n + j) * B + b);
        P[i][j] = v;
        P[j][i] = v;
      }
    }
    nll = ws_load(sg.ws + (n + n * n) * B + b);
    t = ws_load(sg.ws + (n + n * n + 2) * B + (b >> 5));
  }
#pragma unroll U
  for (int i = 0; i < n; ++i) eps[i] = 0.0;
  bool obs_fresh = false;
#pragma unroll
  for (int k = 0; k < NP; ++k) th[k] = a.theta ? a.theta[k * B + b] : a.theta_shared[k];

  const double h = a.h;
  const long long si = a.save_interval;
  if (si > 0 && sg.first) {
    if constexpr (SQ != 0) {
      double Pc[n][n];
      factor_to_cov<n>(P, Pc);
      save_slot<n>(0, B, b, L, x, eps, Pc, false, a.out_x, a.out_eps, a.out_P, a.out_yhat, a.out_S);
      save_factor_slot<n>(0, B, b, P, a.out_Ps);
    } else
    save_slot<n>(0, B, b, L, x, eps, P, false, a.out_x, a.out_eps, a.out_P, a.out_yhat, a.out_S);
    if (b == 0 && a.out_t) a.out_t[0] = t;
  }
  long long next_save = si;  // step count at which the next slot is written
  long long slot = 1;
  // The final-state y_hat / S are those of the LAST measurement update of the run (the reference
  // state dict simply keeps them, sqrt_ekf.py:370-372).  Publishing them at every observation step
  // cost L + L^2 global stores per step (12 of the 12.1 stores per step at C2, +2.5 % run time), so
  // only the step that is the last observation writes them.  The dynamic scheduler finds that step
  // once per launch (find_last_obs_kernel); a static launch scans the flags backwards once per
  // trajectory (one load when every step is observed).
  long long last_obs = -1;
  if (LK != 0 && a.has_obs && (a.yhatT || a.ST)) {
    if (sg.last_obs) {
#ifdef __CUDA_ARCH__
      last_obs = __ldg(sg.last_obs);
#else
      last_obs = *sg.last_obs;
#endif
    } else {
      for (long long s = a.T - 1; s >= sg.step0; --s)
        if (a.flags[s]) { last_obs = s; break; }
    }
  }

  for (long long step = sg.step0; step < sg.step1; ++step) {
#ifdef __CUDA_ARCH__
    if constexpr (STG && LK > 0) {       // chunk boundary: next chunk into flight, then wait for this one
      const long long rel = step - sg.step0;
      if ((rel & (OBS_CH - 1)) == 0) {
        const int k = (int)(rel / OBS_CH) & 1;
        const long long b0 = b - (threadIdx.x & 31);
        if (rel == 0) stage_issue<LK>(a, sg, step, 0, b0);
        if (step + OBS_CH < sg.step1) stage_issue<LK>(a, sg, step + OBS_CH, k ^ 1, b0);
        mbar_wait(sg.stg->bar + k, (*sg.stg->parity >> k) & 1u);
        __syncwarp();
        if ((threadIdx.x & 31) == 0) *sg.stg->parity ^= (1u << k);
        __syncwarp();
      }
    }
#endif
    // ---- predict (src/filters/sqrt_ekf.py:92-197)
    double xn[n], J[n][n];
    if (LK == -1 && a.skip_predict) {
      // single-step FilterCorrect: leave t, x, eps, P untouched
    } else {
    if constexpr (is_implicit<Tab>::value) {
      // implicit solver plugin (dirk.cuh): the whole step Jacobian in one call; the factor form needs
      // T = J P_sqrt
      double Jd[n][n];
      dirk_step_generic<Ode, Tab, double>(t, h, x, th, xn, eps, Jd);
      if constexpr (SQ != 0) {
        for (int i = 0; i < n; ++i)
          for (int k = 0; k < n; ++k) {
            double s_ = 0.0;
            for (int j = 0; j < n; ++j) s_ = fma(Jd[i][j], P[j][k], s_);
            J[i][k] = s_;
          }
      } else {
        for (int i = 0; i < n; ++i)
          for (int k = 0; k < n; ++k) J[i][k] = Jd[i][k];
      }
    } else if constexpr (SQ != 0) {
      // tangents seeded with the columns of P_sqrt like jmp_aux (src/utils.py:72-79): J holds T = J P_sqrt
      static_assert(SQ == 0 || KC == n, "factor form carries all tangent columns in one pass");
      rk_step_tangent<Ode, Tab, KC>(t, h, a.st, x, th, 0, true, xn, eps, J, P);
    } else if constexpr (KC == n) {
      rk_step_tangent<Ode, Tab, KC>(t, h, a.st, x, th, 0, true, xn, eps, J);
    } else {
#pragma unroll 1
      for (int c0 = 0; c0 < n; c0 += KC) {
        double Jc[n][KC];
        rk_step_tangent<Ode, Tab, KC>(t, h, a.st, x, th, c0, c0 == 0, xn, eps, Jc);
#pragma unroll 1
        for (int i = 0; i < n; ++i)
#pragma unroll
          for (int k = 0; k < KC; ++k)
            if (c0 + k < n) J[i][c0 + k] = Jc[i][k];
      }
    }
    if constexpr (SQ == 1) {
      double e[n];
#pragma unroll
      for (int i = 0; i < n; ++i) e[i] = (a.cov_fn == COV_STATIC_DIAGONAL) ? cov_scale : cov_scale * eps[i];
      predict_factor_diag<n>(J, e, P);
    } else if constexpr (SQ == 2) {
      predict_factor_generic<n>(a.noise_mode, a.cov_fn, cov_scale, eps, a.GQs, J, P);
    } else {
    propagate_cov<n>(J, P);
    add_process_noise<n>(a.noise_mode, a.cov_fn, cov_scale, eps, a.GQ, P);
    }
#pragma unroll U
    for (int i = 0; i < n; ++i) x[i] = xn[i];
    t = t + h;  // accumulated like rksolver.py:145 (stage times depend on it)
    }

    // ---- correct + log-likelihood (src/filters/sqrt_ekf.py:337-376, src/utils.py:109-128)
    ObsSink sink;
    sink.stride = B;
    sink.y1 = sink.S1 = sink.y2 = sink.S2 = nullptr;
    if (si > 0 && next_save <= a.T) {                      // a further save point exists
      if (a.out_yhat) sink.y1 = a.out_yhat + slot * L * B + b;
      if (a.out_S) sink.S1 = a.out_S + slot * L * L * B + b;
    }
    if (step == last_obs) {
      if (a.yhatT) sink.y2 = a.yhatT + b;
      if (a.ST) sink.S2 = a.ST + b;
    }
    sink.any = sink.y1 || sink.S1 || sink.y2 || sink.S2;
    if constexpr (LK == -1) {
      if (a.has_obs && a.flags[step]) {
        const long long oi = a.ymap[step];
        double y[n];
#pragma unroll U
        for (int l = 0; l < n; ++l)
          if (l < L) y[l] = a.ys_per_traj ? a.ys[(oi * L + l) * B + b] : a.ys[oi * L + l];
        if constexpr (SQ != 0)
          nll += nll_term(correct_factor_generic<n>(L, a.H, a.Rs, y, x, P, sink, a.guard_verbatim != 0, gc), a.nan_to_num);
        else
        nll += nll_term(correct_step<n>(L, a.H, a.R, y, x, P, sink), a.nan_to_num);
        obs_fresh = true;
      }
    } else if constexpr (LK > 0) {
      if (a.flags[step]) {
        const long long oi = a.ymap[step];
        const double* yp = a.ys_per_traj ? a.ys + oi * LK * B + b : a.ys + oi * LK;   // one address, one stride
        const long long ystr = a.ys_per_traj ? B : 1;
        double y[LK];
#ifdef __CUDA_ARCH__
        if constexpr (STG) {
          const long long rel = step - sg.step0;
          const double* ysm = sg.stg->buf + ((((int)(rel / OBS_CH) & 1) * OBS_CH + (int)(rel & (OBS_CH - 1))) * LK) * 32 + (threadIdx.x & 31);
#pragma unroll
          for (int l = 0; l < LK; ++l) y[l] = ysm[l * 32];
        } else
#endif
        {
#pragma unroll
        for (int l = 0; l < LK; ++l) y[l] = yp[l * ystr];
        }
        if constexpr (SQ == 1)
          nll += nll_term(correct_factor_lead<n, LK>(a.R, a.Rs, y, x, P, sink, a.nan_to_num ? nullptr : &lp,
                                                     a.guard_verbatim != 0, gc), a.nan_to_num);
        else
        nll += nll_term(correct_step_lead<n, LK>(a.R, y, x, P, sink, a.nan_to_num ? nullptr : &lp), a.nan_to_num);
        obs_fresh = true;
      }
    }

    // ---- strided save (scripts/run_filter.py:219-222)
    if (si > 0 && step + 1 == next_save) {
      if constexpr (SQ != 0) {
        double Pc[n][n];
        factor_to_cov<n>(P, Pc);
        save_slot<n>(slot, B, b, L, x, eps, Pc, obs_fresh, a.out_x, a.out_eps, a.out_P, a.out_yhat, a.out_S);
        save_factor_slot<n>(slot, B, b, P, a.out_Ps);
      } else
      save_slot<n>(slot, B, b, L, x, eps, P, obs_fresh, a.out_x, a.out_eps, a.out_P, a.out_yhat, a.out_S);
      obs_fresh = false;
      if (b == 0 && a.out_t) a.out_t[slot] = t;
      ++slot;
      next_save += si;
    }
  }

  if (!sg.last) {   // hand the state to whoever runs the next segment
#pragma unroll U
    for (int i = 0; i < n; ++i) sg.ws[i * B + b] = x[i];
    if constexpr (SQ != 0) {
#pragma unroll U
      for (int i = 0; i < n; ++i)
#pragma unroll U
        for (int j = 0; j < n; ++j) sg.ws[(n + i * n + j) * B + b] = P[i][j];
      sg.ws[(n + n * n + 1) * B + b] =
          LogProd::from_bits(((long long)gc.mismatch << 32) | (long long)(unsigned)gc.fired);
    } else {
#pragma unroll U
    for (int i = 0; i < n; ++i)
#pragma unroll U
      for (int j = 0; j <= i; ++j) sg.ws[(n + i * n + j) * B + b] = P[i][j];
    }
    sg.ws[(n + n * n) * B + b] = nll + lp.flush();
    if ((b & 31) == 0) sg.ws[(n + n * n + 2) * B + (b >> 5)] = t;
    return;
  }
  // ---- final state
#pragma unroll U
  for (int i = 0; i < n; ++i) {
    if (a.xT) a.xT[i * B + b] = x[i];
    if (a.epsT) a.epsT[i * B + b] = eps[i];
  }
  if constexpr (SQ != 0) {
    if (a.PsT) {
#pragma unroll U
      for (int i = 0; i < n; ++i)
#pragma unroll U
        for (int j = 0; j < n; ++j) a.PsT[(i * n + j) * B + b] = P[i][j];
    }
    if (a.PT) {
      double Pc[n][n];
      factor_to_cov<n>(P, Pc);
#pragma unroll U
      for (int i = 0; i < n; ++i)
#pragma unroll U
        for (int j = 0; j < n; ++j) a.PT[(i * n + j) * B + b] = Pc[i][j];
    }
    if (a.guard_counts) {
      a.guard_counts[b] = gc.fired;
      a.guard_counts[B + b] = gc.mismatch;
    }
  } else
  if (a.PT) {
#pragma unroll U
    for (int i = 0; i < n; ++i)
#pragma unroll U
      for (int j = 0; j < n; ++j) a.PT[(i * n + j) * B + b] = P[i][j];
  }
  if (a.nll) a.nll[b] = nll + lp.flush();
  if (b == 0 && a.tT) a.tT[0] = t;
}

template <class Ode, class Tab, int KC, int LK, int BLOCK, int MINB, int SQ = 0>
__global__ void __launch_bounds__(BLOCK, MINB)
ekf_thread_kernel(const __grid_constant__ EkfArgs<Ode::NX, Ode::NP> a) {
  const long long b = (long long)blockIdx.x * BLOCK + threadIdx.x;
  if (b >= a.B) return;
  const Segment whole = {0, a.T, true, true, nullptr, nullptr, nullptr};
  ekf_trajectory<Ode, Tab, KC, LK, SQ>(a, b, whole);
}

// last step that carries an observation (-1: none): one thread, scanned once per run instead of by
// every thread of every time segment
static __global__ void find_last_obs_kernel(const unsigned char* flags, long long T, long long* out) {
  long long last = -1;
  for (long long s = T - 1; s >= 0; --s)
    if (flags[s]) { last = s; break; }
  *out = last;
}

// ---------------------------------------------------------------------------------------------
// Persistent, dynamically scheduled variant (throughput runs, save_interval == 0).
//
// Why: B = 65,536 trajectories are 2,048 warps, i.e. 3.46 per SM sub-partition; with one static
// launch, sub-partitions holding 4 warps set the run time while those holding 3 idle for the
// last quarter (ncu: smsp__inst_executed max/min = 4:3, profiles/r1a_*).  Here the run is cut
// into (block of 32 trajectories) x (time segment) work items handed out through one global
// counter in segment-major order; a warp that draws (block, s) waits until segment s-1 of that
// block has been published (done[block] >= s), picks the state up from the workspace, runs the
// segment and publishes it.  Every item only ever waits on an item that was handed out earlier
// to a warp that is already running, so the scheme cannot deadlock whatever the residency.
struct SchedArgs {
  long long seg_len, nseg, nblk;
  double* ws;                // state workspace (see Segment)
  int* counter;              // next work item
  int* done;                 // [nblk] number of published segments per block
  const long long* last_obs; // nullable (no observations): see Segment
};

template <class Ode, class Tab, int KC, int LK, int BLOCK, int MINB, int SQ = 0, bool STG = false>
__global__ void __launch_bounds__(BLOCK, MINB)
ekf_thread_sched_kernel(const __grid_constant__ EkfArgs<Ode::NX, Ode::NP> a,
                        const __grid_constant__ SchedArgs s) {
  const int lane = threadIdx.x & 31;
  // per-warp observation ring (STG): 2 x OBS_CH x LK lines of 256 bytes + 2 mbarriers
  constexpr int LKS = (STG && LK > 0) ? LK : 1;
  __shared__ __align__(128) double obs_ring[STG ? BLOCK / 32 : 1][STG ? 2 * OBS_CH * LKS * 32 : 1];
  __shared__ unsigned long long obs_bar[STG ? BLOCK / 32 : 1][2];
  __shared__ unsigned obs_par[STG ? BLOCK / 32 : 1];
  ObsStage stage = {nullptr, nullptr, nullptr};
  if constexpr (STG) {
    const int w = threadIdx.x >> 5;
    if (lane == 0) {
      mbar_init(&obs_bar[w][0], 1);
      mbar_init(&obs_bar[w][1], 1);
      obs_par[w] = 0u;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    stage.buf = obs_ring[w];
    stage.bar = obs_bar[w];
    stage.parity = &obs_par[w];
  }
  const long long total = s.nblk * s.nseg;
  for (;;) {
    long long item = 0;
    if (lane == 0) item = atomicAdd(s.counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total) return;
    const long long seg = item / s.nblk;
    const long long blk = item % s.nblk;
    if (seg > 0) {
      if (lane == 0) {
        volatile int* d = s.done + blk;
        while (*d < (int)seg) __nanosleep(64);
      }
      __syncwarp();
      __threadfence();
    }
    const long long b = blk * 32 + lane;
    if (b < a.B) {
      Segment sg;
      sg.step0 = seg * s.seg_len;
      sg.step1 = (seg + 1 == s.nseg) ? a.T : (seg + 1) * s.seg_len;
      sg.first = seg == 0;
      sg.last = seg + 1 == s.nseg;
      sg.ws = s.ws;
      sg.last_obs = s.last_obs;
      sg.stg = STG ? &stage : nullptr;
      ekf_trajectory<Ode, Tab, KC, LK, SQ, STG>(a, b, sg);
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(s.done + blk, (int)seg + 1);
  }
}

}  // namespace odeu
