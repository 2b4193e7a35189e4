// Host launcher of the forward-mode gradient kernel (ekf_grad.cuh).
#pragma once
#include "ekf_rows.cuh"
#include "sens.cuh"
#include <cstdlib>
#include "plan.h"

namespace odeu {

template <class Ode>
struct GradCfg {
  static constexpr int KC = (Ode::NX <= 3) ? Ode::NX : ((Ode::NX == 4) ? 2 : 1);  // state columns / pass
  static constexpr int PC = (Ode::NX <= 3) ? 2 : 1;                               // directions / thread
  static constexpr int BLOCK = 64;
};

// `g` may be null: filter-only use of the row kernel (p_opt = 0).
template <class Ode>
int fill_grad_args(const odeu_plan& plan, const odeu_ekf_io& io, const odeu_grad_io* gp,
                   GradArgs<Ode::NX, Ode::NP>& a) {
  static const odeu_grad_io none = {0, nullptr, nullptr, nullptr};
  const odeu_grad_io& g = gp ? *gp : none;
  constexpr int n = Ode::NX;
  constexpr int NP = Ode::NP;
  if (io.B <= 0 || io.T < 0) { set_error("odeu_ekf_grad_run: B must be > 0 and T >= 0"); return -1; }
  if (io.L < 0 || io.L > n) { set_error("odeu_ekf_grad_run: L=%d outside [0, n=%d]", io.L, n); return -1; }
  if (!io.x0 || (!io.P0_sqrt && !io.P0)) { set_error("odeu_ekf_grad_run: x0 and P0_sqrt (or P0) are required"); return -1; }
  if (io.P0 && gp) { set_error("odeu_ekf_grad_run: per-trajectory P0 is not supported"); return -1; }
  if (gp && (io.cov_scale_batch || io.nll_nan_to_num)) { set_error("odeu_ekf_grad_run: the calibration-sweep options are served by odeu_ekf_run"); return -1; }
  if (gp && (g.p_opt < 1 || g.p_opt > ODEU_MAX_GRAD || !g.idx || !g.grad)) {
    set_error("odeu_ekf_grad_run: need 1..%d parameter indices and a grad buffer", ODEU_MAX_GRAD);
    return -1;
  }
  if (io.L > 0 && (!io.H || !io.R_sqrt || !io.ys || !io.correct_flags || !io.xy_index_map)) {
    set_error("odeu_ekf_grad_run: L > 0 needs H, R_sqrt, ys, correct_flags, xy_index_map");
    return -1;
  }
  a.B = io.B; a.T = io.T; a.t0 = io.t0; a.h = plan.desc.step_size; a.L = io.L;
  a.cov_fn = plan.desc.cov_fn_id; a.cov_scale = plan.desc.cov_scale;
  a.ys_per_traj = io.ys_per_trajectory; a.has_obs = io.L > 0 ? 1 : 0;
  a.p_opt = g.p_opt;
  for (int j = 0; j < ODEU_MAX_GRAD; ++j) a.idx[j] = -1;
  for (int j = 0; j < g.p_opt; ++j) {
    if (g.idx[j] < 0 || g.idx[j] >= NP) { set_error("odeu_ekf_grad_run: parameter index %d out of range", g.idx[j]); return -1; }
    a.idx[j] = g.idx[j];
  }
  a.x0 = io.x0; a.x0_tan = g.x0_tangent; a.theta = io.theta; a.ys = io.ys;
  a.qdiag = io.Q_sqrt_diag_batch; a.qdiag_tan = gp ? g.Q_sqrt_diag_tangent : nullptr; a.q_gamma = io.gamma_sqrt;
  a.flags = io.correct_flags; a.ymap = (const long long*)io.xy_index_map;
  a.nll = io.nll; a.grad = g.grad; a.xT = io.xT;
  // (only the NLL-only row kernel honours the trajectory / final-state outputs below; the eligibility
  // tests keep every other kernel away from runs that ask for them)
  a.save_interval = gp ? 0 : io.save_interval; a.P0b = gp ? nullptr : io.P0;
  a.epsT = io.epsT; a.yhatT = io.yhatT; a.ST = io.ST; a.tT = io.tT;
  a.out_t = io.out_t; a.out_x = io.out_x; a.out_eps = io.out_eps; a.out_P = io.out_P;
  a.out_yhat = io.out_yhat; a.out_S = io.out_S;
  for (int i = 0; i < n * n; ++i) { a.P0s[i] = 0.0; a.GQ[i] = 0.0; a.H[i] = 0.0; a.R[i] = 0.0; }
  if (io.P0_sqrt)
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += io.P0_sqrt[i * n + k] * io.P0_sqrt[j * n + k];
      a.P0s[i * n + j] = s;
    }
  // per-trajectory diag(w): w >= 0 with |w| = sqrt(n), so any(Q_sqrt >= 1e-16) holds (a NaN w, which
  // the reference would route to the no-Q branch, poisons the run here instead)
  bool qany = io.Q_sqrt_diag_batch != nullptr;
  if (io.Q_sqrt && !io.Q_sqrt_diag_batch) {
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
  for (int l = 0; l < io.L; ++l) {
    for (int j = 0; j < n; ++j) a.H[l * n + j] = io.H[l * n + j];
    for (int m = 0; m < io.L; ++m) {
      double s = 0.0;
      for (int k = 0; k < io.L; ++k) s += io.R_sqrt[l * io.L + k] * io.R_sqrt[m * io.L + k];
      a.R[l * io.L + m] = s;
    }
  }
  for (int k = 0; k < NP; ++k)
    a.theta_shared[k] = io.theta_shared ? io.theta_shared[k] : plan.theta_default[k];
  a.h_sel_all = (io.L > 0 && io.L <= 8) ? 1 : 0;
  for (int l = 0; l < 8; ++l) a.h_sel[l] = 0;
  for (int l = 0; l < io.L && l < 8; ++l) {
    int ones = 0, other = 0, at = 0;
    for (int j = 0; j < n; ++j) {
      if (io.H[l * n + j] == 1.0) { ++ones; at = j; }
      else if (io.H[l * n + j] != 0.0) ++other;
    }
    if (ones == 1 && other == 0) a.h_sel[l] = at;
    else a.h_sel_all = 0;
  }
  return 0;
}

template <class Tab, int NXA, int NPA>
void fill_rt_tableau(GradArgs<NXA, NPA>& a) {
  a.rt_S = Tab::S;
  for (int i = 0; i < 8; ++i) {
    a.rt_c[i] = i < Tab::S ? Tab::c(i) : 0.0;
    a.rt_b[0][i] = i < Tab::S ? Tab::b(0, i) : 0.0;
    a.rt_b[1][i] = i < Tab::S ? Tab::b(1, i) : 0.0;
    for (int j = 0; j < 8; ++j) a.rt_A[i][j] = (i < Tab::S && j < Tab::S) ? Tab::a(i, j) : 0.0;
  }
}

// Row-parallel kernel (ekf_rows.cuh) for ODE plugins that expose the row interface.
template <class Ode, class Tab, class S>
constexpr bool rows_static_ok() {
  if constexpr (has_rows<Ode>::value && !is_implicit<Tab>::value) {
    return Ode::NX > 4 && RowsSmem<Ode, Tab, S, 32 / Ode::ROW_GROUPS>::bytes <= 227 * 1024;
  } else {
    return false;
  }
}
template <class Ode, class Tab, class S>
bool rows_eligible(const odeu_ekf_io& io) {
  if constexpr (rows_static_ok<Ode, Tab, S>()) {
    static const bool off = getenv("ODEU_NO_ROWS") != nullptr;   // A/B switch for measurements
    // the NLL-only instantiation (S = double) serves the full output contract of unroll(): strided slots,
    // eps / y_hat / S / t, per-trajectory P0 (resume); the gradient instantiation does not need it
    constexpr bool full = std::is_same<S, double>::value;
    return !off && io.L <= ROWS_LMAX && !io.cov_scale_batch && !io.nll_nan_to_num && !io.skip_predict &&
           io.guard_mode == ODEU_GUARD_INTENDED &&
           (full || (!io.P0 && io.save_interval == 0 && !io.epsT && !io.yhatT && !io.ST && !io.tT));
  } else {
    return false;
  }
}

template <class Ode, class Tab, class S, int LT>
int launch_rows_lt(GradArgs<Ode::NX, Ode::NP>& a, double* PT, cudaStream_t stream) {
  constexpr int TB = 32 / Ode::ROW_GROUPS;
  using SM = RowsSmem<Ode, Tab, S, TB>;
  constexpr int MINB = (2 * (SM::bytes + 1024) <= 228 * 1024) ? 2 : 1;
  auto kern = ekf_rows_kernel<Ode, Tab, S, TB, MINB, LT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::bytes);
  if (e != cudaSuccess) { set_error("row kernel: shared memory opt-in failed: %s", cudaGetErrorString(e)); return (int)e; }
  const long long units = a.B * (a.p_opt > 0 ? a.p_opt : 1);
  constexpr int UPC = TB * lanes_of<S>::value;      // units per CTA
  kern<<<(unsigned)((units + UPC - 1) / UPC), 32 * Ode::ROW_CLASSES, SM::bytes, stream>>>(a, PT);
  return 0;
}

template <class Ode, class Tab, class S>
int launch_rows(GradArgs<Ode::NX, Ode::NP>& a, double* PT, cudaStream_t stream) {
  if constexpr (rows_static_ok<Ode, Tab, S>()) {
    fill_rt_tableau<Tab>(a);
    fill_rows_schedule<Tab>(a);
    // Experiment kept behind ODEU_ROWS_2WIDE=1: two trajectories per thread (V2d scalar), one CTA of
    // 32 trajectories per SM.  Measured on B200 (tools/bench_c3.py, B = 4,096): 25.4 ms per 1,000
    // steps against 22.1 ms for two co-resident one-wide CTAs - the second dependency chain does
    // fill issue slots (a lone 2-wide CTA needs 1.6x, not 2x, the time of a lone 1-wide one), but
    // 255 registers with 1.5 KB of spills lose to plain co-residency.
    if constexpr (std::is_same<S, double>::value && std::is_same<Tab, TabRKF45>::value &&
                  rows_static_ok<Ode, Tab, V2d>()) {
      static const bool two_wide = getenv("ODEU_ROWS_2WIDE") != nullptr;
      if (two_wide && a.save_interval == 0 && !a.P0b && !a.epsT && !a.yhatT && !a.ST && !a.tT) {
        if (a.L == Ode::ROW_GROUPS && a.has_obs) return launch_rows_lt<Ode, Tab, V2d, Ode::ROW_GROUPS>(a, PT, stream);
        return launch_rows_lt<Ode, Tab, V2d, 0>(a, PT, stream);
      }
    }
    // the headline tableau gets the observation dimension at compile time (L = number of
    // compartments / 1: the measurement matrices of configs/params/hodgkinhuxley*.yaml)
    if constexpr (std::is_same<Tab, TabRKF45>::value) {
      if (a.L == Ode::ROW_GROUPS && a.has_obs) return launch_rows_lt<Ode, Tab, S, Ode::ROW_GROUPS>(a, PT, stream);
    }
    return launch_rows_lt<Ode, Tab, S, 0>(a, PT, stream);
  } else {
    return -2;
  }
}

template <class Ode, class Tab>
int launch_sens(const odeu_plan& plan, const odeu_sens_io& s, cudaStream_t stream) {
  constexpr int NP = Ode::NP;
  if (s.B <= 0 || !s.x0 || !s.w) { set_error("odeu_param_sensitivity: B > 0, x0 and w are required"); return -1; }
  if (s.p_opt < 1 || s.p_opt > ODEU_MAX_GRAD || !s.idx) {
    set_error("odeu_param_sensitivity: need 1..%d parameter indices", ODEU_MAX_GRAD);
    return -1;
  }
  SensArgs<NP> a;
  a.B = s.B; a.t0 = s.t0; a.h = plan.desc.step_size; a.p_opt = s.p_opt;
  for (int j = 0; j < ODEU_MAX_GRAD; ++j) a.idx[j] = -1;
  for (int j = 0; j < s.p_opt; ++j) {
    if (s.idx[j] < 0 || s.idx[j] >= NP) { set_error("odeu_param_sensitivity: parameter index %d out of range", s.idx[j]); return -1; }
    a.idx[j] = s.idx[j];
  }
  a.x0 = s.x0; a.x0_tan = s.x0_tangent; a.theta = s.theta; a.w = s.w; a.w_tan = s.w_tangent;
  for (int k = 0; k < NP; ++k) a.theta_shared[k] = s.theta_shared ? s.theta_shared[k] : plan.theta_default[k];
  dim3 grid((unsigned)((s.B + 63) / 64), (unsigned)(s.w_tangent ? s.p_opt : 1));
  param_sens_kernel<Ode, Tab><<<grid, 64, 0, stream>>>(a);
  count_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("odeu_param_sensitivity: launch failed: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

template <class Ode, class Tab>
int launch_grad(const odeu_plan& plan, const odeu_ekf_io* iop, const odeu_grad_io* gp, const odeu_sens_io* sens,
                cudaStream_t stream) {
  if (sens) return launch_sens<Ode, Tab>(plan, *sens, stream);
  const odeu_ekf_io& io = *iop;
  const odeu_grad_io& g = *gp;
  using Cfg = GradCfg<Ode>;
  GradArgs<Ode::NX, Ode::NP> a;
  if (int rc = fill_grad_args<Ode>(plan, io, &g, a)) return rc;
  if constexpr (!is_implicit<Tab>::value)       // (implicit solver plugins: thread-per-unit kernel below)
  if (rows_eligible<Ode, Tab, GDual<double, 1>>(io)) {
    if (int rc = launch_rows<Ode, Tab, GDual<double, 1>>(a, nullptr, stream)) return rc;
    count_launch();
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("odeu_ekf_grad_run: launch failed: %s", cudaGetErrorString(err)); return (int)err; }
    return 0;
  }
  const int nchunks = (g.p_opt + Cfg::PC - 1) / Cfg::PC;
  dim3 grid((unsigned)((io.B + Cfg::BLOCK - 1) / Cfg::BLOCK), (unsigned)nchunks);
  ekf_grad_kernel<Ode, Tab, Cfg::KC, Cfg::PC, Cfg::BLOCK><<<grid, Cfg::BLOCK, 0, stream>>>(a);
  count_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("odeu_ekf_grad_run: launch failed: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

// Filter run of a medium-size system (the Hodgkin-Huxley family) through the row kernel, called by odeu_ekf_run
// first; returns -100 when the run is not eligible (it then takes the thread-per-trajectory kernel).
template <class Ode, class Tab>
int launch_rows_nll(const odeu_plan& plan, const odeu_ekf_io& io, cudaStream_t stream) {
  if (!rows_eligible<Ode, Tab, double>(io)) return -100;
  GradArgs<Ode::NX, Ode::NP> a;
  if (int rc = fill_grad_args<Ode>(plan, io, nullptr, a)) return rc;
  if (int rc = launch_rows<Ode, Tab, double>(a, io.PT, stream)) return rc;
  count_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("odeu_ekf_run: launch failed: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

template <class Ode>
RowsLaunchFn resolve_rows_solver(int solver) {
  switch (solver) {
    case ODEU_SOLVER_RKF45: return &launch_rows_nll<Ode, TabRKF45>;
    case ODEU_SOLVER_DOPRI65: return &launch_rows_nll<Ode, TabDopri65>;
    case ODEU_SOLVER_BS32: return &launch_rows_nll<Ode, TabBS32>;
    case ODEU_SOLVER_HEUN_EULER: return &launch_rows_nll<Ode, TabHeunEuler>;
    default: return nullptr;
  }
}

template <class Ode>
GradLaunchFn resolve_grad_solver(int solver) {
  switch (solver) {
    case ODEU_SOLVER_RKF45: return &launch_grad<Ode, TabRKF45>;
    case ODEU_SOLVER_DOPRI65: return &launch_grad<Ode, TabDopri65>;
    case ODEU_SOLVER_BS32: return &launch_grad<Ode, TabBS32>;
    case ODEU_SOLVER_HEUN_EULER: return &launch_grad<Ode, TabHeunEuler>;
    case ODEU_SOLVER_KVAERNO3: return &launch_grad<Ode, TabKvaerno3>;
    case ODEU_SOLVER_IMPLICIT_EULER: return &launch_grad<Ode, TabImplicitEuler>;
    default: return nullptr;
  }
}

}  // namespace odeu
