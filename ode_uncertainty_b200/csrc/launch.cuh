// Host launcher template: validates an odeu_ekf_io, folds the small shared matrices into the
// by-value kernel arguments and launches the whole-trajectory kernel on the caller's stream.
#pragma once
#include "ekf_thread.cuh"
#include "pf_thread.cuh"
#include <cstdlib>

#include "plan.h"

namespace odeu {

// per-ODE launch shape: BLOCK threads, MINB resident blocks per SM (register cap), KC tangent
// columns carried per RK pass
template <class Ode>
#ifndef ODEU_THREAD_MINB
#define ODEU_THREAD_MINB 6      // 64-thread CTAs per SM for n <= 4: 6 -> up to 168 registers, no spills (measured: Lorenz 18.96 (8) / 21.26 (6) / 20.83 (5) / 20.29 (4) G trajectory-steps/s)
#endif
struct LaunchCfg {
  static constexpr int BLOCK = 64;
  static constexpr int MINB = (Ode::NX <= 4) ? ODEU_THREAD_MINB : 2;
  static constexpr int KC = (Ode::NX <= 4) ? Ode::NX : 1;
};

// Validates `io` and folds the shared host matrices into the by-value kernel arguments.
template <class Tab>
void fill_scaled_tableau(double h, ScaledTableau& st) {
  for (int i = 0; i < 8; ++i) {
    st.hb1[i] = (i < Tab::S) ? h * Tab::b(1, i) : 0.0;
    st.b0[i] = (i < Tab::S) ? Tab::b(0, i) : 0.0;
    st.b1[i] = (i < Tab::S) ? Tab::b(1, i) : 0.0;
    for (int j = 0; j < 8; ++j) {
      st.ha[i][j] = (i < Tab::S && j < Tab::S) ? h * Tab::a(i, j) : 0.0;
      st.a[i][j] = (i < Tab::S && j < Tab::S) ? Tab::a(i, j) : 0.0;
    }
  }
}

template <class Ode>
int fill_ekf_args(const odeu_plan& plan, const odeu_ekf_io& io, EkfArgs<Ode::NX, Ode::NP>& a) {
  constexpr int n = Ode::NX;
  constexpr int NP = Ode::NP;
  if (io.B <= 0 || io.T < 0) { set_error("odeu_ekf_run: B must be > 0 and T >= 0"); return -1; }
  if (io.L < 0 || io.L > n) { set_error("odeu_ekf_run: L=%d outside [0, n=%d]", io.L, n); return -1; }
  if (!io.x0) { set_error("odeu_ekf_run: x0 is required"); return -1; }
  if (!io.P0 && !io.P0_sqrt && !io.P0_sqrt_batch) { set_error("odeu_ekf_run: P0, P0_sqrt or P0_sqrt_batch is required"); return -1; }
  if (io.guard_mode < ODEU_GUARD_INTENDED || io.guard_mode > ODEU_GUARD_INTENDED_FACTOR) {
    set_error("odeu_ekf_run: unknown guard_mode %d", io.guard_mode);
    return -1;
  }
  const bool factor = io.guard_mode != ODEU_GUARD_INTENDED;
  if (factor && n > 4) {
    set_error("odeu_ekf_run: guard_mode reference (factor form) is served for state dimension <= 4, this plan has n = %d", n);
    return -2;
  }
  if (factor && io.P0 && !io.P0_sqrt_batch) {
    set_error("odeu_ekf_run: guard_mode reference resumes from the FACTOR (P0_sqrt_batch = a previous PT_sqrt), not from P0");
    return -1;
  }
  if (!factor && (io.P0_sqrt_batch || io.PT_sqrt || io.out_P_sqrt || io.guard_counts)) {
    set_error("odeu_ekf_run: P0_sqrt_batch / PT_sqrt / out_P_sqrt / guard_counts need guard_mode reference (factor form)");
    return -1;
  }

  if (io.save_interval < 0) { set_error("odeu_ekf_run: save_interval < 0"); return -1; }
  if (io.Q_sqrt_diag_batch) {
    set_error("odeu_ekf_run: the per-trajectory diagonal Q_sqrt (parameter_sensitivity) is served by odeu_ekf_grad_run");
    return -1;
  }
  if (io.L > 0 && (!io.H || !io.R_sqrt || !io.ys || !io.correct_flags || !io.xy_index_map)) {
    set_error("odeu_ekf_run: L > 0 needs H, R_sqrt, ys, correct_flags, xy_index_map");
    return -1;
  }

  a.B = io.B; a.T = io.T; a.t0 = io.t0; a.h = plan.desc.step_size; a.L = io.L;
  a.cov_fn = plan.desc.cov_fn_id; a.cov_scale = plan.desc.cov_scale;
  a.save_interval = io.save_interval;
  a.ys_per_traj = io.ys_per_trajectory; a.has_obs = io.L > 0 ? 1 : 0;
  a.skip_predict = io.skip_predict;
  a.x0 = io.x0; a.P0 = io.P0; a.theta = io.theta; a.ys = io.ys;
  a.flags = io.correct_flags; a.ymap = (const long long*)io.xy_index_map;
  a.xT = io.xT; a.epsT = io.epsT; a.PT = io.PT; a.yhatT = io.yhatT; a.ST = io.ST; a.nll = io.nll;
  a.scale_b = io.cov_scale_batch; a.nan_to_num = io.nll_nan_to_num;
  a.tT = io.tT;
  a.guard_verbatim = io.guard_mode == ODEU_GUARD_REFERENCE ? 1 : 0;
  a.P0f_b = io.P0_sqrt_batch; a.PsT = io.PT_sqrt; a.guard_counts = (long long*)io.guard_counts;
  a.out_Ps = io.save_interval > 0 ? io.out_P_sqrt : nullptr;
  a.out_t = io.out_t; a.out_x = io.out_x; a.out_eps = io.out_eps; a.out_P = io.out_P;
  a.out_yhat = io.out_yhat; a.out_S = io.out_S;

  // P0 = P0_sqrt P0_sqrt^T (scripts/run_filter.py:74-78 builds the factor)
  for (int i = 0; i < n * n; ++i) { a.P0s[i] = 0.0; a.GQ[i] = 0.0; a.H[i] = 0.0; a.R[i] = 0.0; }
  if (io.P0_sqrt)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += io.P0_sqrt[i * n + k] * io.P0_sqrt[j * n + k];
        a.P0s[i * n + j] = s;
      }
  // Qany = any(Q_sqrt >= 1e-16) (sqrt_ekf.py:108,128); GQ = (g Q_sqrt)(g Q_sqrt)^T
  bool qany = false;
  if (io.Q_sqrt) {
    for (int i = 0; i < n * n; ++i) qany = qany || (io.Q_sqrt[i] >= 1e-16);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int k = 0; k < n; ++k)
          s += (io.gamma_sqrt * io.Q_sqrt[i * n + k]) * (io.gamma_sqrt * io.Q_sqrt[j * n + k]);
        a.GQ[i * n + j] = s;
      }
  }
  a.noise_mode = plan.desc.disable_cov_update ? (qany ? NOISE_Q_ONLY : NOISE_NONE)
                                              : (qany ? NOISE_EPS_PLUS_Q : NOISE_COVFN);
  const int L = io.L;
  for (int l = 0; l < L; ++l) {
    for (int j = 0; j < n; ++j) a.H[l * n + j] = io.H[l * n + j];
    for (int m = 0; m < L; ++m) {
      double s = 0.0;
      for (int k = 0; k < L; ++k) s += io.R_sqrt[l * L + k] * io.R_sqrt[m * L + k];
      a.R[l * L + m] = s;
    }
  }
  for (int k = 0; k < NP; ++k)
    a.theta_shared[k] = io.theta_shared ? io.theta_shared[k] : plan.theta_default[k];
  if constexpr (n <= 4) {    // factor form: the factors themselves
    for (int i = 0; i < n * n; ++i) {
      a.P0f[i] = io.P0_sqrt ? io.P0_sqrt[i] : 0.0;
      a.GQs[i] = io.Q_sqrt ? io.gamma_sqrt * io.Q_sqrt[i] : 0.0;
      a.Rs[i] = 0.0;
    }
    for (int i = 0; i < L * L; ++i) a.Rs[i] = io.R_sqrt[i];
  }
  return 0;
}

// Which compile-time measurement-update variant serves this run (see ekf_trajectory's LK):
// small systems get 0 (prediction only), 1 and n when H is the leading identity block
// (every measurement_matrix the reference ships for them), everything else the generic code.
template <class Ode>
int select_lk(const odeu_ekf_io& io) {
  constexpr int n = Ode::NX;
  if (n > 4 || io.skip_predict) return -1;
  if (io.L == 0) return 0;
  if (io.L != 1 && io.L != n) return -1;
  for (int l = 0; l < io.L; ++l)
    for (int j = 0; j < n; ++j)
      if (io.H[l * n + j] != ((l == j) ? 1.0 : 0.0)) return -1;
  return io.L;
}

// ---- dynamic (block, time-segment) scheduling: geometry shared by the size query and the launch
struct SchedGeom {
  long long seg_len, nseg, nblk, state_doubles, bytes;
};
inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}
inline bool sched_geometry(int n, long long B, long long T, SchedGeom& g) {
  g.nblk = (B + 31) / 32;
  static const int nseg_target = getenv("ODEU_SCHED_NSEG") ? atoi(getenv("ODEU_SCHED_NSEG")) : 96;  // tuning knob (measured: 24 < 48 < 96)
  if (nseg_target <= 0) return false;
  g.seg_len = (T + nseg_target - 1) / nseg_target;
  if (g.seg_len < 64) g.seg_len = 64;
  g.nseg = (T + g.seg_len - 1) / g.seg_len;
  g.state_doubles = (long long)(n + n * n + 2) * B + g.nblk + 1;   // + 1: last observation step
  g.bytes = g.state_doubles * 8 + (1 + g.nblk) * 4 + 64;
  // worth it only when every SM sub-partition holds several warps and there are several segments
  return n <= 4 && g.nseg >= 4 && g.nblk >= 4LL * sm_count();
}

template <class Ode, class Tab, int LK, int SQ = 0>
int launch_ekf_variant(const EkfArgs<Ode::NX, Ode::NP>& a, const odeu_ekf_io& io, cudaStream_t stream) {
  using Cfg = LaunchCfg<Ode>;
  SchedGeom g;
  // measured on B200 (tools/sched_sweep.py): +7..12 % with the measurement update (Lorenz 16.1 ->
  // 17.6, Van der Pol 32.4 -> 36.3 G trajectory-steps/s); prediction-only runs are faster with the
  // static launch (32.1 vs 27.3), so LK == 0 keeps it
  if constexpr (LK != 0 && SQ != 2)     // (the generic factor form keeps the static launch)
  if (io.workspace && io.save_interval == 0 && !io.skip_predict &&
      sched_geometry(Ode::NX, a.B, a.T, g) && io.workspace_bytes >= g.bytes) {
    SchedArgs s;
    s.seg_len = g.seg_len; s.nseg = g.nseg; s.nblk = g.nblk;
    s.ws = (double*)io.workspace;
    s.counter = (int*)((char*)io.workspace + g.state_doubles * 8);
    s.done = s.counter + 1;
    cudaError_t e = cudaMemsetAsync(s.counter, 0, (1 + g.nblk) * 4, stream);
    if (e != cudaSuccess) { set_error("odeu_ekf_run: workspace memset failed: %s", cudaGetErrorString(e)); return (int)e; }
    long long* last_obs = (long long*)io.workspace + (g.state_doubles - 1);
    s.last_obs = nullptr;
    if (a.has_obs && (a.yhatT || a.ST)) {
      find_last_obs_kernel<<<1, 1, 0, stream>>>(a.flags, a.T, last_obs);
      s.last_obs = last_obs;
    }
    long long resident = (long long)sm_count() * Cfg::MINB;     // one CTA per register slot
    static const int cap = getenv("ODEU_SCHED_CAP") ? atoi(getenv("ODEU_SCHED_CAP")) : 1;       // tuning knob
    const long long useful = (g.nblk * 32 + Cfg::BLOCK - 1) / Cfg::BLOCK;  // more warps than blocks only spin
    if (cap && resident > useful) resident = useful;
    // per-trajectory observation streams are staged through shared memory by bulk-async copies
    // (whole 256-byte lines per warp: B % 32 == 0; ODEU_NO_OBS_STAGE=1 keeps the plain loads, for A/B runs)
    if constexpr (LK > 0) {
      static const bool no_stage = getenv("ODEU_NO_OBS_STAGE") != nullptr;
      if (a.ys_per_traj && a.B % 32 == 0 && !no_stage) {
        ekf_thread_sched_kernel<Ode, Tab, Cfg::KC, LK, Cfg::BLOCK, Cfg::MINB, SQ, true>
            <<<(unsigned)resident, Cfg::BLOCK, 0, stream>>>(a, s);
        return 0;
      }
    }
    ekf_thread_sched_kernel<Ode, Tab, Cfg::KC, LK, Cfg::BLOCK, Cfg::MINB, SQ>
        <<<(unsigned)resident, Cfg::BLOCK, 0, stream>>>(a, s);
    return 0;
  }
  const long long grid = (a.B + Cfg::BLOCK - 1) / Cfg::BLOCK;
  ekf_thread_kernel<Ode, Tab, Cfg::KC, LK, Cfg::BLOCK, Cfg::MINB, SQ>
      <<<(unsigned)grid, Cfg::BLOCK, 0, stream>>>(a);
  return 0;
}

// Factor form (guard_mode reference): the structured variant needs H = [I_L 0] (LK > 0), the
// NOISE_COVFN branch with a diagonal process-noise block and a diagonal R_sqrt (what every script of
// the reference builds: const_diag(L, sqrt(obs_noise_var))); everything else takes the generic code.
template <class Ode>
bool factor_fast_ok(const EkfArgs<Ode::NX, Ode::NP>& a, const odeu_ekf_io& io, int lk) {
  if (lk <= 0) return false;
  if (a.noise_mode != NOISE_COVFN || a.cov_fn == COV_OUTER) return false;
  for (int l = 0; l < io.L; ++l)
    for (int m = 0; m < io.L; ++m)
      if (m != l && io.R_sqrt[l * io.L + m] != 0.0) return false;
  return true;
}

template <class Ode, class Tab>
int launch_ekf(const odeu_plan& plan, const odeu_ekf_io& io, cudaStream_t stream) {
  constexpr int n = Ode::NX;
  EkfArgs<Ode::NX, Ode::NP> a;
  if (int rc = fill_ekf_args<Ode>(plan, io, a)) return rc;
  fill_scaled_tableau<Tab>(plan.desc.step_size, a.st);
  const int lk = select_lk<Ode>(io);
  int rc = 0;
  if constexpr (is_implicit<Tab>::value) {      // implicit solver plugins: the generic variants only
    if constexpr (n <= 4) {
      if (io.guard_mode != ODEU_GUARD_INTENDED) rc = launch_ekf_variant<Ode, Tab, -1, 2>(a, io, stream);
      else rc = launch_ekf_variant<Ode, Tab, -1>(a, io, stream);
    } else {
      rc = launch_ekf_variant<Ode, Tab, -1>(a, io, stream);
    }
  } else
  if constexpr (n <= 4) {
    if (io.guard_mode != ODEU_GUARD_INTENDED) {
      if (factor_fast_ok<Ode>(a, io, lk)) {
        if (lk == 1) rc = launch_ekf_variant<Ode, Tab, 1, 1>(a, io, stream);
        else rc = launch_ekf_variant<Ode, Tab, n, 1>(a, io, stream);
      } else {
        rc = launch_ekf_variant<Ode, Tab, -1, 2>(a, io, stream);
      }
    } else
    if (lk == 0) rc = launch_ekf_variant<Ode, Tab, 0>(a, io, stream);
    else if (lk == 1) rc = launch_ekf_variant<Ode, Tab, 1>(a, io, stream);
    else if (lk == n) rc = launch_ekf_variant<Ode, Tab, n>(a, io, stream);
    else rc = launch_ekf_variant<Ode, Tab, -1>(a, io, stream);
  } else {
    rc = launch_ekf_variant<Ode, Tab, -1>(a, io, stream);
  }
  if (rc) return rc;
  count_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error("odeu_ekf_run: launch failed: %s", cudaGetErrorString(err));
    return (int)err;
  }
  return 0;
}

template <class Ode>
int fill_pf_args(const odeu_plan& plan, const odeu_pf_io& io, PfArgs<Ode::NX, Ode::NP>& a) {
  constexpr int n = Ode::NX;
  constexpr int NP = Ode::NP;
  if (io.M <= 0 || io.T < 0) { set_error("odeu_pf_run: M must be > 0 and T >= 0"); return -1; }
  if (!io.x0 && !io.x0_shared) { set_error("odeu_pf_run: x0 or x0_shared is required"); return -1; }
  if (io.save_interval < 0) { set_error("odeu_pf_run: save_interval < 0"); return -1; }
  a.M = io.M; a.T = io.T; a.t0 = io.t0; a.h = plan.desc.step_size;
  a.cov_fn = plan.desc.cov_fn_id; a.cov_scale = plan.desc.cov_scale;
  a.seed = io.seed; a.particle_offset = io.particle_offset; a.step_offset = io.step_offset;
  a.save_interval = io.save_interval;
  a.noise_free = io.noise_free;
  a.x0 = io.x0; a.xT = io.xT; a.epsT = io.epsT; a.tT = io.tT;
  a.out_t = io.out_t; a.out_x = io.out_x; a.out_eps = io.out_eps;
  for (int i = 0; i < n; ++i) a.x0s[i] = io.x0_shared ? io.x0_shared[i] : 0.0;
  for (int k = 0; k < NP; ++k)
    a.theta_shared[k] = io.theta_shared ? io.theta_shared[k] : plan.theta_default[k];
  return 0;
}

template <class Ode, class Tab>
int launch_pf(const odeu_plan& plan, const odeu_pf_io& io, cudaStream_t stream) {
  constexpr int BLOCK = 128;
  PfArgs<Ode::NX, Ode::NP> a;
  if (int rc = fill_pf_args<Ode>(plan, io, a)) return rc;
  if constexpr (!is_implicit<Tab>::value) fill_scaled_tableau<Tab>(plan.desc.step_size, a.st);
  const long long grid = (io.M + BLOCK - 1) / BLOCK;
  pf_thread_kernel<Ode, Tab, BLOCK><<<(unsigned)grid, BLOCK, 0, stream>>>(a);
  count_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error("odeu_pf_run: launch failed: %s", cudaGetErrorString(err));
    return (int)err;
  }
  return 0;
}

template <class Ode>
int launch_rhs(const odeu_plan& plan, long long B, double t, const double* x, const double* theta,
               const double* theta_shared, double* dx, cudaStream_t stream) {
  if (B <= 0 || !x || !dx) { set_error("odeu_ode_rhs: invalid argument"); return -1; }
  RhsArgs<Ode::NP> a;
  a.B = B; a.t = t; a.x = x; a.theta = theta; a.dx = dx;
  for (int k = 0; k < Ode::NP; ++k) a.theta_shared[k] = theta_shared ? theta_shared[k] : plan.theta_default[k];
  ode_rhs_kernel<Ode><<<(unsigned)((B + 127) / 128), 128, 0, stream>>>(a);
  count_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("odeu_ode_rhs: launch failed: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

template <class Ode>
Launchers resolve_solver(int solver) {
  switch (solver) {
    case ODEU_SOLVER_RKF45: return {&launch_ekf<Ode, TabRKF45>, &launch_pf<Ode, TabRKF45>, &launch_rhs<Ode>};
    case ODEU_SOLVER_DOPRI65: return {&launch_ekf<Ode, TabDopri65>, &launch_pf<Ode, TabDopri65>, &launch_rhs<Ode>};
    case ODEU_SOLVER_BS32: return {&launch_ekf<Ode, TabBS32>, &launch_pf<Ode, TabBS32>, &launch_rhs<Ode>};
    case ODEU_SOLVER_HEUN_EULER: return {&launch_ekf<Ode, TabHeunEuler>, &launch_pf<Ode, TabHeunEuler>, &launch_rhs<Ode>};
    case ODEU_SOLVER_KVAERNO3: return {&launch_ekf<Ode, TabKvaerno3>, &launch_pf<Ode, TabKvaerno3>, &launch_rhs<Ode>};
    case ODEU_SOLVER_IMPLICIT_EULER: return {&launch_ekf<Ode, TabImplicitEuler>, &launch_pf<Ode, TabImplicitEuler>, &launch_rhs<Ode>};
    default: return {nullptr, nullptr, nullptr};
  }
}

}  // namespace odeu
