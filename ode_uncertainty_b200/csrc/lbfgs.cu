// Lock-step projected L-BFGS over R independent restarts, on the device.
//
// The reference minimises the EKF negative log-likelihood over the normalised parameters in
// [0, 1]^p with SciPy's L-BFGS-B, one OS process per random restart
// (scripts/run_parameter_estimation.py:599-667, p_umap :265-272); src/utils.py:15-36 also carries a
// projected L-BFGS in JAX (update, then projection onto the box).  Here every restart is a small
// state machine advanced by ONE thread; all restarts are advanced by one launch of
// `lbfgs_step_kernel` after every batched objective evaluation (odeu_ekf_grad_run serves the value
// and the forward-mode gradient of all restarts in one launch), so an optimiser iteration never
// leaves the GPU: no device-to-host read, no per-restart host thread.
//
// Per restart and evaluation (z = trial point just evaluated, f, g = its value and gradient):
//   line search   Armijo backtracking on the PROJECTED path z(alpha) = clip(z_k + alpha d, 0, 1):
//                 accept when f <= f_k + c1 g_k . (z - z_k), else alpha <- alpha / 2 and try again;
//   on accept     curvature pair s = z - z_k, y = g - g_k (kept when s.y > 1e-10 |s||y|, history of
//                 m = 10), convergence tests (projected-gradient max-norm <= pgtol, relative decrease
//                 <= ftol, iteration cap: SciPy's defaults pgtol = 1e-5, factr = 1e7), then the next
//                 direction by the two-loop recursion on the free variables (a variable is held when
//                 it sits on a bound and the gradient pushes outward), steepest descent as fallback
//                 when the quasi-Newton direction is not a descent direction;
//   next trial    z_k + alpha d projected onto the box, alpha = 1 (first iteration: 1 / |g|).
// A restart that rejects a trial does not wait for the others: every launch advances every restart by
// exactly one evaluation, like the independent SciPy runs would.
#include "plan.h"

namespace odeu {

constexpr int LB_M = 10;        // history length (SciPy L-BFGS-B default maxcor)
constexpr int LB_PMAX = 32;     // ODEU_MAX_GRAD

struct LbfgsArgs {
  int R, p, maxiter, first;
  double pgtol, ftol, c1;
  // per restart (device): current iterate and bookkeeping
  double* z;        // [R][p] accepted iterate
  double* f;        // [R]
  double* g;        // [R][p] gradient at z (w.r.t. the normalised parameters)
  double* d;        // [R][p] search direction
  double* S;        // [R][M][p]
  double* Y;        // [R][M][p]
  double* rho;      // [R][M]
  double* alpha;    // [R]
  int* meta;        // [R][6]: history count, history head, iterations, evaluations, status, ls trials
  // the trial point that was just evaluated, its value and gradient (physical-parameter gradient
  // scaled by `scale` = max - min -> normalised)
  double* zt;       // [R][p] in: evaluated trial, out: next trial
  const double* ft; // [R]
  const double* gt; // [R][p]
  const double* scale;  // [p]
};

enum { LB_RUNNING = 0, LB_CONV_PG = 1, LB_CONV_F = 2, LB_MAXITER = 3, LB_LS_FAIL = 4, LB_NAN = 5 };

__device__ inline double clip01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }

__global__ void __launch_bounds__(64) lbfgs_step_kernel(const __grid_constant__ LbfgsArgs a) {
  const int r = blockIdx.x * 64 + threadIdx.x;
  if (r >= a.R) return;
  const int p = a.p;
  double* z = a.z + (size_t)r * p;
  double* g = a.g + (size_t)r * p;
  double* d = a.d + (size_t)r * p;
  double* zt = a.zt + (size_t)r * p;
  double* Sm = a.S + (size_t)r * LB_M * p;
  double* Ym = a.Y + (size_t)r * LB_M * p;
  double* rho = a.rho + (size_t)r * LB_M;
  int* meta = a.meta + r * 6;
  if (meta[4] != LB_RUNNING) return;            // finished restarts keep their trial point = iterate
  double gn[LB_PMAX];
  for (int i = 0; i < p; ++i) gn[i] = a.gt[(size_t)r * p + i] * a.scale[i];
  const double ft = a.ft[r];
  meta[3] += 1;
  bool accept;
  if (a.first) {
    accept = true;
  } else {
    double dec = 0.0;
    for (int i = 0; i < p; ++i) dec += g[i] * (zt[i] - z[i]);
    accept = ft <= a.f[r] + a.c1 * dec;         // false for NaN
  }
  if (!(ft == ft) && a.first) { meta[4] = LB_NAN; return; }
  if (!accept) {
    meta[5] += 1;
    const double al = a.alpha[r] * 0.5;
    a.alpha[r] = al;
    if (meta[5] >= 30) { meta[4] = LB_LS_FAIL; for (int i = 0; i < p; ++i) zt[i] = z[i]; return; }
    for (int i = 0; i < p; ++i) zt[i] = clip01(z[i] + al * d[i]);
    return;
  }
  // ---- accepted: history update, convergence tests
  double f_old = a.f[r];
  if (!a.first) {
    double sy = 0.0, ss = 0.0, yy = 0.0;
    int head = meta[1];
    double* sn = Sm + (size_t)head * p;
    double* yn = Ym + (size_t)head * p;
    for (int i = 0; i < p; ++i) {
      const double s = zt[i] - z[i], y = gn[i] - g[i];
      sn[i] = s; yn[i] = y;
      sy += s * y; ss += s * s; yy += y * y;
    }
    if (sy > 1e-10 * sqrt(ss * yy) && sy > 0.0) {
      rho[head] = 1.0 / sy;
      meta[1] = (head + 1) % LB_M;
      if (meta[0] < LB_M) meta[0] += 1;
    }
    meta[2] += 1;
  }
  for (int i = 0; i < p; ++i) { z[i] = zt[i]; g[i] = gn[i]; }
  a.f[r] = ft;
  meta[5] = 0;
  // free variables / projected gradient
  bool held[LB_PMAX];
  double pg = 0.0;
  for (int i = 0; i < p; ++i) {
    held[i] = (z[i] <= 0.0 && gn[i] > 0.0) || (z[i] >= 1.0 && gn[i] < 0.0);
    const double pgi = z[i] - clip01(z[i] - gn[i]);      // SciPy's projected gradient
    pg = fmax(pg, fabs(pgi));
  }
  if (pg <= a.pgtol) { meta[4] = LB_CONV_PG; return; }
  if (!a.first && (f_old - ft) <= a.ftol * fmax(fmax(fabs(f_old), fabs(ft)), 1.0)) { meta[4] = LB_CONV_F; return; }
  if (meta[2] >= a.maxiter) { meta[4] = LB_MAXITER; return; }
  // ---- two-loop recursion on the free variables
  double q[LB_PMAX], al[LB_M];
  for (int i = 0; i < p; ++i) q[i] = held[i] ? 0.0 : gn[i];
  const int cnt = meta[0], head = meta[1];
  for (int k = 0; k < cnt; ++k) {
    const int j = (head - 1 - k + 2 * LB_M) % LB_M;
    const double* sj = Sm + (size_t)j * p;
    const double* yj = Ym + (size_t)j * p;
    double s = 0.0;
    for (int i = 0; i < p; ++i) s += (held[i] ? 0.0 : sj[i]) * q[i];
    al[k] = rho[j] * s;
    for (int i = 0; i < p; ++i) q[i] -= al[k] * (held[i] ? 0.0 : yj[i]);
  }
  if (cnt > 0) {
    const int j = (head - 1 + LB_M) % LB_M;
    const double* sj = Sm + (size_t)j * p;
    const double* yj = Ym + (size_t)j * p;
    double sy = 0.0, yy = 0.0;
    for (int i = 0; i < p; ++i) { sy += sj[i] * yj[i]; yy += yj[i] * yj[i]; }
    const double gamma = sy / yy;
    for (int i = 0; i < p; ++i) q[i] *= gamma;
  }
  for (int k = cnt - 1; k >= 0; --k) {
    const int j = (head - 1 - k + 2 * LB_M) % LB_M;
    const double* sj = Sm + (size_t)j * p;
    const double* yj = Ym + (size_t)j * p;
    double s = 0.0;
    for (int i = 0; i < p; ++i) s += (held[i] ? 0.0 : yj[i]) * q[i];
    const double be = rho[j] * s;
    for (int i = 0; i < p; ++i) q[i] += (al[k] - be) * (held[i] ? 0.0 : sj[i]);
  }
  double dg = 0.0, gg = 0.0;
  for (int i = 0; i < p; ++i) { d[i] = held[i] ? 0.0 : -q[i]; dg += d[i] * gn[i]; gg += (held[i] ? 0.0 : gn[i] * gn[i]); }
  double alpha = 1.0;
  if (!(dg < 0.0) || cnt == 0) {                 // not a descent direction / no curvature yet: steepest descent
    for (int i = 0; i < p; ++i) d[i] = held[i] ? 0.0 : -gn[i];
    alpha = cnt == 0 ? fmin(1.0, 1.0 / sqrt(gg)) : 1.0;
  }
  a.alpha[r] = alpha;
  for (int i = 0; i < p; ++i) zt[i] = clip01(z[i] + alpha * d[i]);
}

}  // namespace odeu

extern "C" int64_t odeu_lbfgs_workspace_doubles(int32_t R, int32_t p) {
  // z, g, d [R][p]; f, alpha [R]; S, Y [R][M][p]; rho [R][M]; meta [R][6] ints (3 doubles)
  return (int64_t)R * (3 * p + 2 + 2 * odeu::LB_M * p + odeu::LB_M + 3);
}

// Advances every restart by one evaluation.  workspace: odeu_lbfgs_workspace_doubles(R, p) doubles,
// ZEROED before the first call; zt [R][p] holds the points that were evaluated (in) and receives the
// next trial points (out); ft [R], gt [R][p] their values and PHYSICAL gradients; scale [p] = max - min.
// first != 0 on the call that follows the evaluation of the starting points.  status_out [R] (device,
// optional): 0 running, 1 projected gradient <= pgtol, 2 relative decrease <= ftol, 3 iteration cap,
// 4 line search failed, 5 NaN start.
extern "C" int odeu_lbfgs_step(int32_t R, int32_t p, int32_t maxiter, int32_t first, double pgtol, double ftol,
                               double* workspace_dev, double* zt_dev, const double* ft_dev, const double* gt_dev,
                               const double* scale_dev, void* cuda_stream) {
  using namespace odeu;
  if (R <= 0 || p <= 0 || p > LB_PMAX || !workspace_dev || !zt_dev || !ft_dev || !gt_dev || !scale_dev) {
    set_error("odeu_lbfgs_step: invalid argument (1 <= p <= %d)", LB_PMAX);
    return -1;
  }
  LbfgsArgs a;
  a.R = R; a.p = p; a.maxiter = maxiter; a.first = first; a.pgtol = pgtol; a.ftol = ftol; a.c1 = 1e-4;
  double* w = workspace_dev;
  a.z = w; w += (size_t)R * p;
  a.g = w; w += (size_t)R * p;
  a.d = w; w += (size_t)R * p;
  a.f = w; w += R;
  a.alpha = w; w += R;
  a.S = w; w += (size_t)R * LB_M * p;
  a.Y = w; w += (size_t)R * LB_M * p;
  a.rho = w; w += (size_t)R * LB_M;
  a.meta = (int*)w;
  a.zt = zt_dev; a.ft = ft_dev; a.gt = gt_dev; a.scale = scale_dev;
  lbfgs_step_kernel<<<(unsigned)((R + 63) / 64), 64, 0, (cudaStream_t)cuda_stream>>>(a);
  count_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("odeu_lbfgs_step: launch failed: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}
