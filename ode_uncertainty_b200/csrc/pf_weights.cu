// Bootstrap-weight update for the particle ensemble (EXTENSION: the reference's ParticleFilter
// has no weights, no correct step and no resampling - src/filters/particle_filter.py:24-118,
// SURVEY F5 - so this step has no reference oracle; BASELINE config 4 names it).
//   logw_m += log N(y; H x_m, R) = -0.5 d^T R^-1 d - 0.5 log det(2 pi R),  d = y - H x_m
#include "plan.h"

namespace odeu {
struct PfWeightArgs {
  long long M;
  int n, L;
  const double* x;   // [n][M]
  double* logw;      // [M]
  double y[16], H[16 * 16], Rinv[16 * 16];
  double logdet_term;  // -0.5 * (L log 2pi + log det R)
};
__global__ void __launch_bounds__(256) pf_weight_kernel(const __grid_constant__ PfWeightArgs a) {
  const long long m = (long long)blockIdx.x * 256 + threadIdx.x;
  if (m >= a.M) return;
  double d[16];
  for (int l = 0; l < a.L; ++l) {
    double s = 0.0;
    for (int j = 0; j < a.n; ++j) s = fma(a.H[l * a.n + j], a.x[j * a.M + m], s);
    d[l] = a.y[l] - s;
  }
  double q = 0.0;
  for (int l = 0; l < a.L; ++l) {
    double s = 0.0;
    for (int k = 0; k < a.L; ++k) s = fma(a.Rinv[l * a.L + k], d[k], s);
    q = fma(d[l], s, q);
  }
  a.logw[m] += -0.5 * q + a.logdet_term;
}
}  // namespace odeu

extern "C" int odeu_pf_weight_update(int64_t M, int32_t n, int32_t L, const double* x_dev,
                                     const double* y_host, const double* H_host, const double* R_host,
                                     double* logw_dev, void* cuda_stream) {
  using namespace odeu;
  if (M <= 0 || n <= 0 || n > 16 || L <= 0 || L > 16 || !x_dev || !y_host || !H_host || !R_host || !logw_dev) {
    set_error("odeu_pf_weight_update: invalid argument (n, L <= 16)");
    return -1;
  }
  PfWeightArgs a;
  a.M = M; a.n = n; a.L = L; a.x = x_dev; a.logw = logw_dev;
  for (int l = 0; l < L; ++l) a.y[l] = y_host[l];
  for (int i = 0; i < L * n; ++i) a.H[i] = H_host[i];
  // R^-1 and log det R by Cholesky on the host (L <= 16)
  double Lc[16][16] = {{0}};
  double logdet = 0.0;
  for (int j = 0; j < L; ++j) {
    double s = R_host[j * L + j];
    for (int k = 0; k < j; ++k) s -= Lc[j][k] * Lc[j][k];
    if (!(s > 0.0)) { set_error("odeu_pf_weight_update: R is not positive definite"); return -1; }
    Lc[j][j] = sqrt(s);
    logdet += 2.0 * log(Lc[j][j]);
    for (int i = j + 1; i < L; ++i) {
      double v = R_host[i * L + j];
      for (int k = 0; k < j; ++k) v -= Lc[i][k] * Lc[j][k];
      Lc[i][j] = v / Lc[j][j];
    }
  }
  for (int c = 0; c < L; ++c) {   // solve R z = e_c
    double w[16], z[16];
    for (int i = 0; i < L; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) s -= Lc[i][k] * w[k];
      w[i] = s / Lc[i][i];
    }
    for (int i = L - 1; i >= 0; --i) {
      double s = w[i];
      for (int k = i + 1; k < L; ++k) s -= Lc[k][i] * z[k];
      z[i] = s / Lc[i][i];
    }
    for (int i = 0; i < L; ++i) a.Rinv[i * L + c] = z[i];
  }
  a.logdet_term = -0.5 * (L * 1.8378770664093453 + logdet);
  pf_weight_kernel<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(a);
  count_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("odeu_pf_weight_update: launch failed: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}
