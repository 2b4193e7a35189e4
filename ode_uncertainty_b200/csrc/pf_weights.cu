// Bootstrap-weight update for the particle ensemble (EXTENSION: the reference's ParticleFilter
// has no weights, no correct step and no resampling - src/filters/particle_filter.py:24-118,
// SURVEY F5 - so this step has no reference oracle; BASELINE config 4 names it).
//   logw_m += log N(y; H x_m, R) = -0.5 d^T R^-1 d - 0.5 log det(2 pi R),  d = y - H x_m
#include <cstdlib>

#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

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

// ---------------------------------------------------------------------------------------------
// Sync-free global steps of the bootstrap filter (EXTENSION, parity unpinned like the rest of this
// file).  Round 1 ran them as eager torch ops with five host round trips and six collectives per
// observation (299 ms against 15.5 ms of prediction at 8 GPUs).  Now per
// observation, all on one stream and without a single device-to-host read:
//   odeu_pf_weight_reduce   logw += log N(y; Hx, R) fused with this rank's (max, sum exp, sum exp^2)
//   [one all-gather of G x 3 doubles]
//   odeu_pf_normalize       global log-sum-exp from the G triples, logw -= lse, ESS, resampling
//                           decision and running log-likelihood as DEVICE scalars; packs (x, w) rows
//   [one all-gather of the packed particles, cumsum of the weights]
//   odeu_pf_resample        systematic resampling by binary search in the global CDF, predicated on
//                           the device-side decision; writes the kernel layout [n][M] of odeu_pf_run
namespace odeu {

__device__ __forceinline__ void lse_merge(double& m, double& s, double& s2, double m2, double t, double t2) {
  // (m, s, s2) represent sum exp = s e^m, sum exp^2 = s2 e^(2m)
  if (m2 > m) { const double f = exp(m - m2); s = fma(s, f, t); s2 = fma(s2, f * f, t2); m = m2; }
  else if (m2 > -1.0e300) { const double f = exp(m2 - m); s = fma(t, f, s); s2 = fma(t2, f * f, s2); }
}

struct PfReduceArgs {
  PfWeightArgs w;
  double* block_part;    // [nblk][3]
  unsigned* ticket;      // zero on entry; reset by the last block
  double* out;           // [3] this rank's (max, sum exp(lw - max), sum exp(2 (lw - max)))
};

// PF_IPT particles per thread: 4 for large shards (fewer, fatter blocks: 26 instead of 35 us at 10^6 particles), 1 when
// the shard has too few particles to fill the SMs with fat blocks (125,000: 9.5 us against 14.5 us with 4)
template <int PF_IPT>
__global__ void __launch_bounds__(256) pf_weight_reduce_kernel(const __grid_constant__ PfReduceArgs a) {
  const PfWeightArgs& w = a.w;
  // PF_IPT particles per thread (consecutive blocks of 256: coalesced), all loads of a thread in flight together; the
  // per-thread maximum enters the block maximum, then one exp per particle
  double lwv[PF_IPT];
  const long long base = (long long)blockIdx.x * (256 * PF_IPT) + threadIdx.x;
  double lmax = -1.0e308;
#pragma unroll
  for (int it = 0; it < PF_IPT; ++it) {
    const long long m = base + (long long)it * 256;
    double lw = -1.0e308;
    if (m < w.M) {
      double d[16];
      for (int l = 0; l < w.L; ++l) {
        double s = 0.0;
        for (int j = 0; j < w.n; ++j) s = fma(w.H[l * w.n + j], w.x[j * w.M + m], s);
        d[l] = w.y[l] - s;
      }
      double q = 0.0;
      for (int l = 0; l < w.L; ++l) {
        double s = 0.0;
        for (int k = 0; k < w.L; ++k) s = fma(w.Rinv[l * w.L + k], d[k], s);
        q = fma(d[l], s, q);
      }
      lw = w.logw[m] + (-0.5 * q + w.logdet_term);
      w.logw[m] = lw;
    }
    lwv[it] = lw;
    lmax = fmax(lmax, lw);
  }
  // block (max, sum, sum2)
  double mx = lmax;
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  __shared__ double sh[3][8];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[0][wid] = mx;
  __syncthreads();
  mx = sh[0][0];
  for (int k = 1; k < 8; ++k) mx = fmax(mx, sh[0][k]);
  double e = 0.0, e2 = 0.0;
#pragma unroll
  for (int it = 0; it < PF_IPT; ++it) {
    const double ei = (base + (long long)it * 256 < w.M) ? exp(lwv[it] - mx) : 0.0;
    e += ei;
    e2 = fma(ei, ei, e2);
  }
  for (int o = 16; o > 0; o >>= 1) { e += __shfl_xor_sync(0xffffffffu, e, o); e2 += __shfl_xor_sync(0xffffffffu, e2, o); }
  __syncthreads();
  if (lane == 0) { sh[1][wid] = e; sh[2][wid] = e2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0, s2 = 0.0;
    for (int k = 0; k < 8; ++k) { s += sh[1][k]; s2 += sh[2][k]; }     // fixed order: deterministic
    a.block_part[3 * blockIdx.x + 0] = mx;
    a.block_part[3 * blockIdx.x + 1] = s;
    a.block_part[3 * blockIdx.x + 2] = s2;
    __threadfence();
    last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last block folds the block partials, each thread a strided slice, then a fixed-order tree
  double M0 = -1.0e308, S = 0.0, S2 = 0.0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += 256)
    lse_merge(M0, S, S2, __ldcg(a.block_part + 3 * b), __ldcg(a.block_part + 3 * b + 1), __ldcg(a.block_part + 3 * b + 2));
  __shared__ double red[3][256];
  red[0][threadIdx.x] = M0; red[1][threadIdx.x] = S; red[2][threadIdx.x] = S2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      double m1 = red[0][threadIdx.x], s1 = red[1][threadIdx.x], t1 = red[2][threadIdx.x];
      lse_merge(m1, s1, t1, red[0][threadIdx.x + o], red[1][threadIdx.x + o], red[2][threadIdx.x + o]);
      red[0][threadIdx.x] = m1; red[1][threadIdx.x] = s1; red[2][threadIdx.x] = t1;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    a.out[0] = red[0][0]; a.out[1] = red[1][0]; a.out[2] = red[2][0];
    *a.ticket = 0u;
  }
}

struct PfNormArgs {
  long long M, M_total;
  int n, G;
  const double* triples;   // [G][3]
  const double* x;         // [n][M]
  double* logw;            // [M] in/out
  double* pack;            // [M][n + 1] rows (x, w)
  double* stats;           // [4]: lse, ess, resample flag (0/1), running log-likelihood
  double* ess_hist;        // this observation's slot, or null
  double* flag_hist;
  double ess_frac;
  double* w_out;           // nullable: [M] the weights once more, contiguous (what the peers scan)
};

__global__ void __launch_bounds__(256) pf_normalize_kernel(const __grid_constant__ PfNormArgs a) {
  double M0 = -1.0e308, S = 0.0, S2 = 0.0;
  for (int g = 0; g < a.G; ++g) lse_merge(M0, S, S2, a.triples[3 * g], a.triples[3 * g + 1], a.triples[3 * g + 2]);
  const double lse = M0 + log(S);
  const long long m = (long long)blockIdx.x * 256 + threadIdx.x;
  if (m == 0) {
    const double ess = S * S / S2;
    const double flag = ess < a.ess_frac * (double)a.M_total ? 1.0 : 0.0;
    a.stats[0] = lse; a.stats[1] = ess; a.stats[2] = flag; a.stats[3] += lse;
    if (a.ess_hist) *a.ess_hist = ess;
    if (a.flag_hist) *a.flag_hist = flag;
  }
  if (m >= a.M) return;
  const double lw = a.logw[m] - lse;
  a.logw[m] = lw;
  double* row = a.pack + m * (a.n + 1);
  for (int i = 0; i < a.n; ++i) row[i] = a.x[i * a.M + m];
  const double w = exp(lw);
  row[a.n] = w;
  if (a.w_out) a.w_out[m] = w;
}

struct PfResampleArgs {
  long long M, M_total, slot_lo;
  int n;
  double u0;
  const double* stats;     // [2] = flag
  const double* cdf;       // [M_total] inclusive cumulative sum of the gathered weights
  const double* pack;      // [M_total][n + 1] gathered rows
  double* x_new;           // [n][M]
  double* logw;            // [M]
};

__global__ void __launch_bounds__(256) pf_resample_kernel(const __grid_constant__ PfResampleArgs a) {
  if (a.stats[2] == 0.0) return;                       // device-side decision: ESS high enough, keep everything
  const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
  if (j >= a.M) return;
  const double total = a.cdf[a.M_total - 1];
  const double u = ((double)(a.slot_lo + j) + a.u0) / (double)a.M_total * total;
  long long lo = 0, hi = a.M_total - 1;               // first index with cdf >= u (the last particle closes the CDF)
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (a.cdf[mid] >= u) hi = mid; else lo = mid + 1;
  }
  const double* row = a.pack + lo * (a.n + 1);
  for (int i = 0; i < a.n; ++i) a.x_new[i * a.M + j] = row[i];
  a.logw[j] = -log((double)a.M_total);
}

// resampling fused with the "keep" branch: no copy of the ensemble is needed on the host side
struct PfResample2Args {
  PfResampleArgs r;
  const double* x_old;     // [n][M] (the prediction's output)
};
__global__ void __launch_bounds__(256) pf_resample_or_keep_kernel(const __grid_constant__ PfResample2Args b) {
  const PfResampleArgs& a = b.r;
  const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
  if (j >= a.M) return;
  if (a.stats[2] == 0.0) {                             // device-side decision: ESS high enough, keep everything
    for (int i = 0; i < a.n; ++i) a.x_new[i * a.M + j] = b.x_old[i * a.M + j];
    return;
  }
  const double total = a.cdf[a.M_total - 1];
  const double u = ((double)(a.slot_lo + j) + a.u0) / (double)a.M_total * total;
  long long lo = 0, hi = a.M_total - 1;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (a.cdf[mid] >= u) hi = mid; else lo = mid + 1;
  }
  const double* row = a.pack + lo * (a.n + 1);
  for (int i = 0; i < a.n; ++i) a.x_new[i * a.M + j] = row[i];
  a.logw[j] = -log((double)a.M_total);
}
// weight column of the packed rows as a scan input
struct PackWeight {
  const double* pack;
  int stride;
  __host__ __device__ double operator()(long long m) const { return pack[m * stride + stride - 1]; }
};
using PackWeightIt = thrust::transform_iterator<PackWeight, thrust::counting_iterator<long long>, double>;

// ---- peer-memory variant (one process per GPU, buffers mapped into every rank over NVLink / NVSwitch):
// no collective library on the data path.  Each rank keeps its packed rows and weights in its own symmetric
// buffer; the peers read them in place - the CDF scan streams the G weight arrays (8 bytes per particle), the
// resampling fetches ONLY the ancestor rows it needs (one 32-byte sector each at n = 3) - instead of every rank
// receiving the whole ensemble in an all-gather.
constexpr int PF_MAX_PEERS = 16;
struct PfPeers {
  const double* pack[PF_MAX_PEERS];   // rank q's packed rows [M][n + 1]
  const double* w[PF_MAX_PEERS];      // rank q's weights [M]
};
struct PfPublishArgs {
  const double* triple;               // this rank's (max, sum, sum of squares)
  double* dst[PF_MAX_PEERS];          // rank q's triples array [G][3]
  int rank, G;
};
__global__ void pf_publish_triple_kernel(const __grid_constant__ PfPublishArgs a) {
  const int q = threadIdx.x / 3, c = threadIdx.x % 3;
  if (q < a.G) a.dst[q][3 * a.rank + c] = a.triple[c];      // remote stores; the barrier that follows orders them
}
struct PeerWeight {                   // global particle index -> weight, read from the owning rank
  PfPeers p;
  long long M;
  __device__ double operator()(long long m) const {
    const int q = (int)(m / M);
    return p.w[q][m - (long long)q * M];
  }
};
using PeerWeightIt = thrust::transform_iterator<PeerWeight, thrust::counting_iterator<long long>, double>;
// weights of all ranks into one local array (rank q's slice pulled from its owner with wide coalesced loads, many in
// flight per SM: a scan that reads the peers through its input iterator is latency-bound on NVLink, 52 us for 10^6
// weights over 8 GPUs against ~15 us for pull + local scan)
struct PfPullArgs {
  PfPeers p;
  double* dst;             // [G * M]
  long long M;
};
__global__ void __launch_bounds__(256) pf_pull_weights_kernel(const __grid_constant__ PfPullArgs a) {
  const int q = blockIdx.y;
  const double* src = a.p.w[q];
  double* dst = a.dst + (long long)q * a.M;
  const long long stride = (long long)gridDim.x * 256;
  long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (((a.M & 1) == 0) && ((reinterpret_cast<unsigned long long>(src) | reinterpret_cast<unsigned long long>(dst)) & 15ull) == 0) {
    const long long h = a.M >> 1;
    for (; i < h; i += stride) reinterpret_cast<double2*>(dst)[i] = __ldcg(reinterpret_cast<const double2*>(src) + i);
  } else {
    for (; i < a.M; i += stride) dst[i] = __ldcg(src + i);
  }
}
struct PfResamplePeerArgs {
  PfResampleArgs r;                   // (r.pack unused)
  const double* x_old;
  PfPeers p;
};
__global__ void __launch_bounds__(256) pf_resample_peer_kernel(const __grid_constant__ PfResamplePeerArgs b) {
  const PfResampleArgs& a = b.r;
  const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
  if (j >= a.M) return;
  if (a.stats[2] == 0.0) {
    for (int i = 0; i < a.n; ++i) a.x_new[i * a.M + j] = b.x_old[i * a.M + j];
    return;
  }
  const double total = a.cdf[a.M_total - 1];
  const double u = ((double)(a.slot_lo + j) + a.u0) / (double)a.M_total * total;
  long long lo = 0, hi = a.M_total - 1;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (a.cdf[mid] >= u) hi = mid; else lo = mid + 1;
  }
  const int q = (int)(lo / a.M);
  const double* row = b.p.pack[q] + (lo - (long long)q * a.M) * (a.n + 1);     // local or over NVLink
  if (a.n == 3) {                                                               // one 32-byte sector
    const double2 r0 = __ldcg(reinterpret_cast<const double2*>(row));
    const double r2 = __ldcg(row + 2);
    a.x_new[j] = r0.x; a.x_new[a.M + j] = r0.y; a.x_new[2 * a.M + j] = r2;
  } else {
    for (int i = 0; i < a.n; ++i) a.x_new[i * a.M + j] = __ldcg(row + i);
  }
  a.logw[j] = -log((double)a.M_total);
}

static int fill_weight_args(PfWeightArgs& a, int64_t M, int32_t n, int32_t L, const double* x_dev, const double* y_host,
                            const double* H_host, const double* R_host, double* logw_dev) {
  a.M = M; a.n = n; a.L = L; a.x = x_dev; a.logw = logw_dev;
  for (int l = 0; l < L; ++l) a.y[l] = y_host[l];
  for (int i = 0; i < L * n; ++i) a.H[i] = H_host[i];
  double Lc[16][16] = {{0}};
  double logdet = 0.0;
  for (int j = 0; j < L; ++j) {
    double s = R_host[j * L + j];
    for (int k = 0; k < j; ++k) s -= Lc[j][k] * Lc[j][k];
    if (!(s > 0.0)) { set_error("odeu_pf_weight: R is not positive definite"); return -1; }
    Lc[j][j] = sqrt(s);
    logdet += 2.0 * log(Lc[j][j]);
    for (int i = j + 1; i < L; ++i) {
      double v = R_host[i * L + j];
      for (int k = 0; k < j; ++k) v -= Lc[i][k] * Lc[j][k];
      Lc[i][j] = v / Lc[j][j];
    }
  }
  for (int c = 0; c < L; ++c) {
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
  return 0;
}
}  // namespace odeu

extern "C" int64_t odeu_pf_reduce_scratch_bytes(int64_t M) { return ((M + 255) / 256) * 3 * 8 + 64; }   // (an upper bound)

extern "C" int odeu_pf_weight_reduce(int64_t M, int32_t n, int32_t L, const double* x_dev, const double* y_host,
                                     const double* H_host, const double* R_host, double* logw_dev,
                                     double* triple_out_dev, void* scratch_dev, void* cuda_stream) {
  using namespace odeu;
  if (M <= 0 || n <= 0 || n > 16 || L <= 0 || L > 16 || !x_dev || !y_host || !H_host || !R_host || !logw_dev ||
      !triple_out_dev || !scratch_dev) {
    set_error("odeu_pf_weight_reduce: invalid argument (n, L <= 16)");
    return -1;
  }
  PfReduceArgs a;
  if (int rc = fill_weight_args(a.w, M, n, L, x_dev, y_host, H_host, R_host, logw_dev)) return rc;
  a.ticket = (unsigned*)scratch_dev;                 // caller zeroes the scratch once; the kernel re-arms it
  a.block_part = (double*)((char*)scratch_dev + 64);
  a.out = triple_out_dev;
  if (M >= 300000) pf_weight_reduce_kernel<4><<<(unsigned)((M + 1023) / 1024), 256, 0, (cudaStream_t)cuda_stream>>>(a);
  else pf_weight_reduce_kernel<1><<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(a);
  count_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("odeu_pf_weight_reduce: launch failed: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

extern "C" int odeu_pf_normalize(int64_t M, int64_t M_total, int32_t n, int32_t G, const double* triples_dev,
                                 const double* x_dev, double* logw_dev, double* pack_dev, double* stats_dev,
                                 double* ess_hist_dev, double* flag_hist_dev, double ess_frac, void* cuda_stream) {
  using namespace odeu;
  if (M <= 0 || M_total < M || n <= 0 || n > 16 || G <= 0 || !triples_dev || !x_dev || !logw_dev || !pack_dev || !stats_dev) {
    set_error("odeu_pf_normalize: invalid argument");
    return -1;
  }
  PfNormArgs a = {M, M_total, n, G, triples_dev, x_dev, logw_dev, pack_dev, stats_dev, ess_hist_dev, flag_hist_dev, ess_frac, nullptr};
  pf_normalize_kernel<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(a);
  count_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("odeu_pf_normalize: launch failed: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

extern "C" int odeu_pf_resample(int64_t M, int64_t M_total, int64_t slot_lo, int32_t n, double u0,
                                const double* stats_dev, const double* cdf_dev, const double* pack_dev,
                                double* x_new_dev, double* logw_dev, void* cuda_stream) {
  using namespace odeu;
  if (M <= 0 || M_total < M || n <= 0 || n > 16 || !stats_dev || !cdf_dev || !pack_dev || !x_new_dev || !logw_dev) {
    set_error("odeu_pf_resample: invalid argument");
    return -1;
  }
  PfResampleArgs a = {M, M_total, slot_lo, n, u0, stats_dev, cdf_dev, pack_dev, x_new_dev, logw_dev};
  pf_resample_kernel<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(a);
  count_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("odeu_pf_resample: launch failed: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

// Inclusive cumulative sum of the gathered weights (CUB scan straight off the strided weight column of the packed
// rows, no contiguous copy) + systematic resampling predicated on the device-side decision; when the decision is
// "keep", x_old is carried over, so the caller ping-pongs two buffers and never copies the ensemble.
extern "C" int64_t odeu_pf_scan_bytes(int64_t M_total) {
  using namespace odeu;
  size_t tmp = 0;
  PackWeightIt it(thrust::counting_iterator<long long>(0), PackWeight{nullptr, 2});
  cub::DeviceScan::InclusiveSum(nullptr, tmp, it, (double*)nullptr, (long long)M_total);
  return (int64_t)(M_total * 8 + ((tmp + 255) / 256) * 256 + 256);
}

extern "C" int odeu_pf_scan_resample(int64_t M, int64_t M_total, int64_t slot_lo, int32_t n, double u0,
                                     const double* stats_dev, const double* pack_dev, const double* x_old_dev,
                                     double* x_new_dev, double* logw_dev, void* scan_dev, int64_t scan_bytes,
                                     void* cuda_stream) {
  using namespace odeu;
  if (M <= 0 || M_total < M || n <= 0 || n > 16 || !stats_dev || !pack_dev || !x_old_dev || !x_new_dev || !logw_dev ||
      !scan_dev || scan_bytes < odeu_pf_scan_bytes(M_total)) {
    set_error("odeu_pf_scan_resample: invalid argument (scan scratch: odeu_pf_scan_bytes(M_total) bytes)");
    return -1;
  }
  cudaStream_t st = (cudaStream_t)cuda_stream;
  double* cdf = (double*)scan_dev;
  void* tmp = (char*)scan_dev + ((M_total * 8 + 255) / 256) * 256;
  size_t tmp_bytes = (size_t)scan_bytes - (size_t)((M_total * 8 + 255) / 256) * 256;
  PackWeightIt it(thrust::counting_iterator<long long>(0), PackWeight{pack_dev, n + 1});
  cudaError_t err = cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, it, cdf, (long long)M_total, st);
  if (err != cudaSuccess) { set_error("odeu_pf_scan_resample: scan failed: %s", cudaGetErrorString(err)); return (int)err; }
  count_launch();
  PfResample2Args a = {{M, M_total, slot_lo, n, u0, stats_dev, cdf, pack_dev, x_new_dev, logw_dev}, x_old_dev};
  pf_resample_or_keep_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(a);
  count_launch();
  err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("odeu_pf_scan_resample: launch failed: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

// ---- peer-memory entry points (see PfPeers).  Pointer tables are HOST arrays of G device pointers, one per rank,
// as torch.distributed._symmetric_memory (or cudaIpcOpenMemHandle) hands them out; G <= 16.
extern "C" int odeu_pf_publish_triple(const double* triple_dev, double* const* peer_triples_host, int32_t rank, int32_t G,
                                      void* cuda_stream) {
  using namespace odeu;
  if (!triple_dev || !peer_triples_host || G <= 0 || G > PF_MAX_PEERS || rank < 0 || rank >= G) {
    set_error("odeu_pf_publish_triple: invalid argument (G <= 16)");
    return -1;
  }
  PfPublishArgs a;
  a.triple = triple_dev; a.rank = rank; a.G = G;
  for (int q = 0; q < PF_MAX_PEERS; ++q) a.dst[q] = q < G ? peer_triples_host[q] : nullptr;
  pf_publish_triple_kernel<<<1, 3 * PF_MAX_PEERS, 0, (cudaStream_t)cuda_stream>>>(a);
  count_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("odeu_pf_publish_triple: launch failed: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

extern "C" int odeu_pf_normalize_w(int64_t M, int64_t M_total, int32_t n, int32_t G, const double* triples_dev,
                                   const double* x_dev, double* logw_dev, double* pack_dev, double* w_dev,
                                   double* stats_dev, double* ess_hist_dev, double* flag_hist_dev, double ess_frac,
                                   void* cuda_stream) {
  using namespace odeu;
  if (M <= 0 || M_total < M || n <= 0 || n > 16 || G <= 0 || !triples_dev || !x_dev || !logw_dev || !pack_dev || !w_dev ||
      !stats_dev) {
    set_error("odeu_pf_normalize_w: invalid argument");
    return -1;
  }
  PfNormArgs a = {M, M_total, n, G, triples_dev, x_dev, logw_dev, pack_dev, stats_dev, ess_hist_dev, flag_hist_dev, ess_frac, w_dev};
  pf_normalize_kernel<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(a);
  count_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("odeu_pf_normalize_w: launch failed: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

extern "C" int odeu_pf_scan_resample_peer(int64_t M, int64_t M_total, int64_t slot_lo, int32_t n, int32_t G, double u0,
                                          const double* stats_dev, const double* const* peer_pack_host,
                                          const double* const* peer_w_host, const double* x_old_dev, double* x_new_dev,
                                          double* logw_dev, void* scan_dev, int64_t scan_bytes, void* cuda_stream) {
  using namespace odeu;
  if (M <= 0 || M_total != M * G || n <= 0 || n > 16 || G <= 0 || G > PF_MAX_PEERS || !stats_dev || !peer_pack_host ||
      !peer_w_host || !x_old_dev || !x_new_dev || !logw_dev || !scan_dev || scan_bytes < odeu_pf_scan_bytes(M_total)) {
    set_error("odeu_pf_scan_resample_peer: invalid argument (M_total = G M, G <= 16, scan scratch: odeu_pf_scan_bytes)");
    return -1;
  }
  cudaStream_t st = (cudaStream_t)cuda_stream;
  PfPeers peers;
  for (int q = 0; q < PF_MAX_PEERS; ++q) {
    peers.pack[q] = q < G ? peer_pack_host[q] : nullptr;
    peers.w[q] = q < G ? peer_w_host[q] : nullptr;
  }
  double* cdf = (double*)scan_dev;
  void* tmp = (char*)scan_dev + ((M_total * 8 + 255) / 256) * 256;
  size_t tmp_bytes = (size_t)scan_bytes - (size_t)((M_total * 8 + 255) / 256) * 256;
  static const bool remote_scan = getenv("ODEU_PF_REMOTE_SCAN") != nullptr;      // A/B switch: scan through the peers
  cudaError_t err;
  if (remote_scan) {
    PeerWeightIt it(thrust::counting_iterator<long long>(0), PeerWeight{peers, (long long)M});
    err = cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, it, cdf, (long long)M_total, st);
  } else {
    PfPullArgs pa = {peers, cdf, (long long)M};
    long long bx = (M / 2 + 255) / 256;
    if (bx > 64) bx = 64;
    if (bx < 1) bx = 1;
    pf_pull_weights_kernel<<<dim3((unsigned)bx, (unsigned)G), 256, 0, st>>>(pa);
    count_launch();
    err = cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, cdf, cdf, (long long)M_total, st);     // in place
  }
  if (err != cudaSuccess) { set_error("odeu_pf_scan_resample_peer: scan failed: %s", cudaGetErrorString(err)); return (int)err; }
  count_launch();
  PfResamplePeerArgs a = {{M, M_total, slot_lo, n, u0, stats_dev, cdf, nullptr, x_new_dev, logw_dev}, x_old_dev, peers};
  pf_resample_peer_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(a);
  count_launch();
  err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("odeu_pf_scan_resample_peer: launch failed: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}
