// Large-state EKF path (BASELINE config 5: coupled-oscillator network, n = 2 D = 256 states,
// 1,024 trajectories): here P <- J P J^T + Q is a real dense contraction (2 x 2 n^3 flops per
// trajectory-step, 67 MFLOP at n = 256), so it runs on the FP64 tensor-core instruction
// (DMMA, mma.sync.m8n8k4.f64) with shared-memory tiles staged by cp.async.  tcgen05 has no f64
// kind, so DMMA is the tensor path that keeps the reference's precision (x64).
//
// One EKF step = four launches on the caller's stream:
//   1. lcao_jac_kernel      RK step of the oscillator chain (src/ode/lcao.py:51-61 generalised to
//                           D oscillators; src/solvers/rksolver.py:113-155) + the step Jacobian
//                           J = d x_next / d x.  The chain couples oscillator i only with D-1-i and
//                           with its own velocity, so the four rows {i, D-1-i, D+i, 2D-1-i} of any
//                           tangent column are closed under the stage recursion: thread (column c)
//                           carries them in registers, no exchange at all; J is written once.
//   2. dgemm_nt_kernel      M = J P        (P symmetric: J P = J P^T, so both products are "NT")
//   3. dgemm_nt_kernel      P = M J^T + Q  (Q diagonal: embedded error / static / tempered noise,
//                           src/filters/sqrt_ekf.py:96-136)
//   4. dense_correct_kernel measurement update for a component-selecting H (L <= 16 observed
//                           states): S = P[h,h] + R, Cholesky, K = P[:,h] S^-1, x += K d, NLL term
//                           (src/utils.py:109-128); emits U = [K G], V = [PH^T K], G = PH^T - K S
//   5. dgemm_nt_kernel<UPD> P <- P - U V^T = P - K (HP) - G K^T (Joseph form up to its
//                           rounding-level residual) as a k = 32 DMMA update, HBM-bound
//
// Layout (device, float64): x [B][n], P [B][n][n] row-major per trajectory, observations
// ys [T_obs][L] (shared) or [T_obs][B][L].  Workspace: J, M [B][n][n] and the noise diagonal [B][n].
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>

#include "plan.h"
#include "tableaux.cuh"

namespace odeu {

enum { DN_COVFN_DIAG = 0, DN_COVFN_STATIC = 1, DN_EPS_PLUS_Q = 2, DN_Q_ONLY = 3, DN_NONE = 4 };

struct DenseStepArgs {
  int D, n;
  long long B;
  double h, t;
  double th0, th1, th2;
  int noise_mode;
  double cov_scale;
  const double* gq_diag;   // [n] gamma * diag(Q_sqrt Q_sqrt^T) (device) or null
  double* x;               // [B][n] in/out
  double* eps;             // [B][n] out (nullable)
  double* J;               // [B][n][n]
  double* qd;              // [B][n] process-noise diagonal of this step
};

// ---------------------------------------------------------------------------------------------
// RK step + Jacobian.  One CTA per trajectory, n threads (thread t: state row t in the primal
// phase, Jacobian column t in the tangent phase).
template <class Tab>
__global__ void __launch_bounds__(512) lcao_jac_kernel(const DenseStepArgs a) {
  constexpr int S = Tab::S;
  extern __shared__ double sm[];
  const int D = a.D, n = a.n;
  double* Xs = sm;             // [n]       stage state
  double* qA = sm + n;         // [S][D]    d f_{D+i} / d x_i at every stage
  const int t = threadIdx.x;
  const long long b = blockIdx.x;
  const double h = a.h;
  const double x0 = a.x[b * n + t];
  double k[S];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    double acc = 0.0;
    bool first = true;
#pragma unroll
    for (int j = 0; j < s; ++j)
      if (Tab::a(s, j) != 0.0) { acc = first ? k[j] * Tab::a(s, j) : fma(Tab::a(s, j), k[j], acc); first = false; }
    const double xi = first ? x0 : fma(h, acc, x0);
    Xs[t] = xi;
    __syncthreads();
    if (t < D) {
      k[s] = Xs[D + t];
      qA[s * D + t] = -a.th0 - 3.0 * a.th1 * (xi * xi);
    } else {
      const int i = t - D;
      const double xp = Xs[i];
      k[s] = (-a.th0) * xp - a.th1 * (xp * xp * xp) - a.th2 * Xs[D - 1 - i];
    }
    __syncthreads();
  }
  {
    double s1 = 0.0, s0 = 0.0;
    bool f1 = true, f0 = true;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      if (Tab::b(1, j) != 0.0) { s1 = f1 ? k[j] * Tab::b(1, j) : fma(Tab::b(1, j), k[j], s1); f1 = false; }
      if (Tab::b(0, j) != 0.0) { s0 = f0 ? k[j] * Tab::b(0, j) : fma(Tab::b(0, j), k[j], s0); f0 = false; }
    }
    const double x1 = fma(h, s1, x0);          // row b[1] propagates (rksolver.py:63-64)
    const double xe = fma(h, s0, x0);
    const double e = fabs(xe - x1);
    a.x[b * n + t] = x1;
    if (a.eps) a.eps[b * n + t] = e;
    double q = 0.0;
    if (a.noise_mode == DN_COVFN_DIAG) q = (a.cov_scale * e) * (a.cov_scale * e);
    else if (a.noise_mode == DN_COVFN_STATIC) q = a.cov_scale * a.cov_scale;
    else if (a.noise_mode == DN_EPS_PLUS_Q) q = e * e + (a.gq_diag ? a.gq_diag[t] : 0.0);
    else if (a.noise_mode == DN_Q_ONLY) q = a.gq_diag ? a.gq_diag[t] : 0.0;
    a.qd[b * n + t] = q;
  }
  // ---- tangent column c = t: rows {i, D-1-i, D+i, 2D-1-i}, i = 0 .. D/2-1
  const int c = t;
  double* Jb = a.J + b * (long long)n * n;
  for (int i = 0; i < D / 2; ++i) {
    const int i2 = D - 1 - i;
    const int rows[4] = {i, i2, D + i, D + i2};
    double e[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) e[m] = (rows[m] == c) ? 1.0 : 0.0;
    double K[S][4];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      double Y[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        double y = e[m];
#pragma unroll
        for (int j = 0; j < s; ++j)
          if (Tab::a(s, j) != 0.0) y = fma(h * Tab::a(s, j), K[j][m], y);
        Y[m] = y;
      }
      K[s][0] = Y[2];
      K[s][1] = Y[3];
      K[s][2] = fma(qA[s * D + i], Y[0], -a.th2 * Y[1]);
      K[s][3] = fma(qA[s * D + i2], Y[1], -a.th2 * Y[0]);
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      double v = e[m];
#pragma unroll
      for (int j = 0; j < S; ++j)
        if (Tab::b(1, j) != 0.0) v = fma(h * Tab::b(1, j), K[j][m], v);
      Jb[(long long)rows[m] * n + c] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Batched C = A B^T (+ diag), all n x n row-major per batch entry, n a multiple of 128.
// CTA tile 128 x 64 (default; 8 warps as 4 x 2, warp tile 32 x 32 = 4 x 4 DMMA m8n8k4 tiles, 2 CTAs
// per SM) or 128 x 128 (2 x 4 warps of 64 x 32); k chunks of 16 double-buffered with cp.async.
// Shared rows are padded to 20 doubles so every fragment load is bank-conflict free.
constexpr int GT = 128, GK = 16, GLD = 20;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& d0, double& d1, double av, double bv) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(av), "d"(bv));
}

// SYM: the product is symmetric (P = M J^T = J P J^T): only tiles on and below the diagonal block
// row are computed, tiles entirely below it are stored twice (the mirrored store writes 64-byte runs).
// CTA tile 128 x BN, WM x WN warps, warp tile (128 / WM) x (BN / WN).
// UPD: C <- C - A B^T with narrow operands (leading dimension ld, inner dimension kdim): the
// rank-2L covariance update of the measurement step, P <- P - [K G] [PH^T K]^T.
template <bool SYM, bool UPD, int BN, int WM, int WN>
__global__ void __launch_bounds__(32 * WM * WN, (BN <= 64 ? 2 : 1))
dgemm_nt_kernel(const double* __restrict__ A, const double* __restrict__ Bm, double* __restrict__ C,
                const double* __restrict__ qd, int n, int ld, int kdim, const unsigned char* __restrict__ flag) {
  if (UPD && flag && !*flag) return;
  const int bm = blockIdx.y * GT, bn = blockIdx.x * BN;
  if (SYM && bn >= bm + GT) return;
  const bool mirror = SYM && (bn + BN <= bm);
  extern __shared__ __align__(16) double gsm[];
  double* As = gsm;                       // [2][GT][GLD]
  double* Bs = gsm + 2 * GT * GLD;        // [2][BN][GLD]
  const long long b = blockIdx.z;
  const double* Ab = A + b * (long long)n * ld + (long long)bm * ld;
  const double* Bb = Bm + b * (long long)n * ld + (long long)bn * ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NTHR = 32 * WM * WN;
  constexpr int MT = GT / (8 * WM), NT = BN / (8 * WN);   // 8-row / 8-column MMA tiles per warp
  const int wm = warp / WN, wn = warp % WN;
  const int lr = lane >> 2, lc = lane & 3;

  double acc[MT][NT][2];
  double* Cb = C + b * (long long)n * n;
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      if (UPD) {   // accumulators start from C (loads overlap the tile copies); A enters negated
        const double2 old = *reinterpret_cast<const double2*>(
            Cb + (long long)(bm + wm * (8 * MT) + i * 8 + lr) * n + bn + wn * (8 * NT) + j * 8 + 2 * lc);
        acc[i][j][0] = old.x;
        acc[i][j][1] = old.y;
      } else {
        acc[i][j][0] = acc[i][j][1] = 0.0;
      }
    }

  auto load_tiles = [&](int buf, int k0) {
#pragma unroll
    for (int q = 0; q < GT * 8 / NTHR; ++q) {
      const int ch = tid + NTHR * q;         // 8 16-byte chunks per row
      const int row = ch >> 3, c2 = (ch & 7) * 2;
      cp_async16(As + ((buf * GT + row) * GLD + c2), Ab + (long long)row * ld + k0 + c2);
    }
#pragma unroll
    for (int q = 0; q < BN * 8 / NTHR; ++q) {
      const int ch = tid + NTHR * q;
      const int row = ch >> 3, c2 = (ch & 7) * 2;
      cp_async16(Bs + ((buf * BN + row) * GLD + c2), Bb + (long long)row * ld + k0 + c2);
    }
    cp_async_commit();
  };

  const int nk = kdim / GK;
  load_tiles(0, 0);
  for (int kc = 0; kc < nk; ++kc) {
    const int buf = kc & 1;
    if (kc + 1 < nk) { load_tiles(buf ^ 1, (kc + 1) * GK); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const double* Aw = As + (buf * GT + wm * (8 * MT) + lr) * GLD + lc;
    const double* Bw = Bs + (buf * BN + wn * (8 * NT) + lr) * GLD + lc;
#pragma unroll
    for (int kk = 0; kk < GK / 4; ++kk) {
      double af[MT], bf[NT];
#pragma unroll
      for (int i = 0; i < MT; ++i) af[i] = UPD ? -Aw[i * 8 * GLD + kk * 4] : Aw[i * 8 * GLD + kk * 4];
#pragma unroll
      for (int j = 0; j < NT; ++j) bf[j] = Bw[j * 8 * GLD + kk * 4];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int row = bm + wm * (8 * MT) + i * 8 + lr;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int col = bn + wn * (8 * NT) + j * 8 + 2 * lc;
      double v0 = acc[i][j][0], v1 = acc[i][j][1];
      if (qd) {
        if (row == col) v0 += qd[b * n + row];
        if (row == col + 1) v1 += qd[b * n + row];
      }
      *reinterpret_cast<double2*>(Cb + (long long)row * n + col) = make_double2(v0, v1);
      if (mirror) {
        Cb[(long long)col * n + row] = v0;
        Cb[(long long)(col + 1) * n + row] = v1;
      }
    }
  }
}

template <int BN, int WM, int WN>
static void launch_gemms(const double* J, double* P, double* M, const double* qd, int n, long long B, cudaStream_t st) {
  const size_t smem = sizeof(double) * 2 * (GT + BN) * GLD;
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(dgemm_nt_kernel<false, false, BN, WM, WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(dgemm_nt_kernel<true, false, BN, WM, WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    done = true;
  }
  const dim3 grid(n / BN, n / GT, (unsigned)B);
  dgemm_nt_kernel<false, false, BN, WM, WN><<<grid, 32 * WM * WN, smem, st>>>(J, P, M, nullptr, n, n, n, nullptr);   // M = J P
  dgemm_nt_kernel<true, false, BN, WM, WN><<<grid, 32 * WM * WN, smem, st>>>(M, J, P, qd, n, n, n, nullptr);          // P = M J^T + Q
}
// P <- P - U V^T, U = [K G], V = [PH^T K] ([B][n][2 DL]); skipped when *flag == 0
static void launch_update(const double* U, const double* V, double* P, int n, long long B, const unsigned char* flag,
                          cudaStream_t st) {
  constexpr int BN = 64, WM = 4, WN = 2;
  const size_t smem = sizeof(double) * 2 * (GT + BN) * GLD;
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(dgemm_nt_kernel<true, true, BN, WM, WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    done = true;
  }
  const dim3 grid(n / BN, n / GT, (unsigned)B);
  dgemm_nt_kernel<true, true, BN, WM, WN><<<grid, 32 * WM * WN, smem, st>>>(U, V, P, nullptr, n, 2 * 16, 2 * 16, flag);
}

// ---------------------------------------------------------------------------------------------
// Measurement update for H = rows of the identity (observed component indices hidx[L], L <= 16).
// One CTA per trajectory, n threads (thread j: column j of P).
constexpr int DL = 16;

struct DenseCorrectArgs {
  int n, L;
  long long B, step;
  int hidx[DL];
  double R[DL * DL];               // R = R_sqrt R_sqrt^T
  const double* ys; int ys_per_traj;
  const unsigned char* flags; const long long* ymap;
  double* x; double* P; double* nll;
  double* U; double* V;            // [B][n][2 DL]: [K G] and [PH^T K] for the rank-2L update GEMM
};

// Step 4a: one WARP per trajectory (lane = observation row): S = P[h,h] + R, Cholesky, innovation,
// NLL term, S^-1 (lane l solves for column l).  The serial chains of the small factorisation run
// in B independent warps side by side instead of stalling a whole CTA per trajectory.
// Outputs per trajectory (workspace): Sw = [S | S^-1 | d] = 2 DL^2 + DL doubles.  With the zero-gain
// guard (see ekf_core.cuh) S^-1 is stored as 0, so K = 0 follows without a branch downstream.
constexpr int SW = 2 * DL * DL + DL;
constexpr int SLD = DL + 1;

__global__ void __launch_bounds__(256) dense_innov_kernel(const DenseCorrectArgs a, double* __restrict__ Sw) {
  if (!a.flags[a.step]) return;
  __shared__ double sh[8][2 * DL * SLD + 2 * DL];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long b = (long long)blockIdx.x * 8 + w;
  if (b >= a.B) return;
  const int n = a.n, L = a.L;
  double* Sm = sh[w];               // [DL][SLD]
  double* Ls = Sm + DL * SLD;       // [DL][SLD]
  double* dv = Ls + DL * SLD;       // [DL]
  double* iv = dv + DL;             // [DL]
  const double* Pb = a.P + b * (long long)n * n;
  const long long oi = a.ymap[a.step];
  const bool act = lane < L;
  if (act) {
    const int hl = a.hidx[lane];
    for (int m = 0; m < L; ++m) Sm[lane * SLD + m] = Pb[(long long)hl * n + a.hidx[m]] + a.R[lane * L + m];
    const double y = a.ys_per_traj ? a.ys[(oi * a.B + b) * L + lane] : a.ys[oi * L + lane];
    dv[lane] = y - a.x[b * n + hl];                              // y_hat = H x
  }
  __syncwarp();
  for (int c = 0; c < L; ++c) {          // Cholesky S = Ls Ls^T, lane = row
    if (lane == c) {
      double s = Sm[c * SLD + c];
      for (int k = 0; k < c; ++k) s = fma(-Ls[c * SLD + k], Ls[c * SLD + k], s);
      const double d = sqrt(s);
      Ls[c * SLD + c] = d;
      iv[c] = 1.0 / d;
    }
    __syncwarp();
    if (lane > c && act) {
      double v = Sm[lane * SLD + c];
      for (int k = 0; k < c; ++k) v = fma(-Ls[lane * SLD + k], Ls[c * SLD + k], v);
      Ls[lane * SLD + c] = v * iv[c];
    }
    __syncwarp();
  }
  // z = Ls^-1 d (column sweep), NLL term (src/utils.py:109-128), zero-gain guard
  double r = act ? dv[lane] : 0.0, z = 0.0;
  for (int c = 0; c < L; ++c) {
    const double zc = __shfl_sync(0xffffffffu, r, c) * iv[c];
    if (lane == c) z = zc;
    if (lane > c && act) r = fma(-Ls[lane * SLD + c], zc, r);
  }
  bool tiny = true;
  double part = 0.0;
  if (act) {
    part = 0.5 * z * z + log(fabs(Ls[lane * SLD + lane]));
    for (int k = 0; k <= lane; ++k) tiny = tiny && (fabs(Ls[lane * SLD + k]) < 1e-16);
  }
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  const bool all_tiny = __all_sync(0xffffffffu, tiny);
  if (lane == 0) a.nll[b] += part + 0.5 * (double)L * 1.8378770664093453;
  // S^-1 column `lane`: Ls Ls^T v = e_lane
  double* out = Sw + b * SW;
  if (act) {
    double wv[DL], v[DL];
    for (int i = 0; i < L; ++i) {
      double s = (i == lane) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) s = fma(-Ls[i * SLD + k], wv[k], s);
      wv[i] = s * iv[i];
    }
    for (int i = L - 1; i >= 0; --i) {
      double s = wv[i];
      for (int k = i + 1; k < L; ++k) s = fma(-Ls[k * SLD + i], v[k], s);
      v[i] = s * iv[i];
    }
    for (int m = 0; m < DL; ++m) {
      out[m * DL + lane] = (m < L) ? Sm[m * SLD + lane] : 0.0;
      out[DL * DL + m * DL + lane] = (m < L && !all_tiny) ? v[m] : 0.0;
    }
    out[2 * DL * DL + lane] = dv[lane];
  } else if (lane < DL) {
    for (int m = 0; m < DL; ++m) { out[m * DL + lane] = 0.0; out[DL * DL + m * DL + lane] = 0.0; }
    out[2 * DL * DL + lane] = 0.0;
  }
}

// Step 4b: thread (trajectory, state j): K_j = (P H^T)_j S^-1, G_j = (P H^T)_j - K_j S, x_j += K_j d;
// emits U = [K G], V = [PH^T K] for the rank-2L update.  No serial section.
__global__ void __launch_bounds__(512) dense_gain_kernel(const DenseCorrectArgs a, const double* __restrict__ Sw) {
  if (!a.flags[a.step]) return;
  __shared__ __align__(16) double sw[SW];
  const int n = a.n, L = a.L;
  const int j = threadIdx.x;
  const long long b = blockIdx.x;
  for (int e = j; e < SW; e += n) sw[e] = Sw[b * SW + e];
  const double* Pb = a.P + b * (long long)n * n;
  double ph[DL];
#pragma unroll
  for (int l = 0; l < DL; ++l) ph[l] = (l < L) ? Pb[(long long)a.hidx[l] * n + j] : 0.0;   // row h_l of the symmetric P
  double xn = a.x[b * n + j];
  __syncthreads();
  const double* Sm = sw;
  const double* Si = sw + DL * DL;
  const double* dv = sw + 2 * DL * DL;
  double kj[DL], g[DL];
#pragma unroll
  for (int l = 0; l < DL; ++l) { kj[l] = 0.0; g[l] = ph[l]; }
#pragma unroll
  for (int m = 0; m < DL; ++m) {
#pragma unroll
    for (int l = 0; l < DL; l += 2) {
      const double2 si = *reinterpret_cast<const double2*>(Si + m * DL + l);
      kj[l] = fma(ph[m], si.x, kj[l]);
      kj[l + 1] = fma(ph[m], si.y, kj[l + 1]);
    }
  }
#pragma unroll
  for (int m = 0; m < DL; ++m) {
#pragma unroll
    for (int l = 0; l < DL; l += 2) {
      const double2 sm2 = *reinterpret_cast<const double2*>(Sm + m * DL + l);
      g[l] = fma(-kj[m], sm2.x, g[l]);
      g[l + 1] = fma(-kj[m], sm2.y, g[l + 1]);
    }
  }
#pragma unroll
  for (int l = 0; l < DL; ++l) xn = fma(kj[l], dv[l], xn);
  a.x[b * n + j] = xn;
  // P <- P - K (HP) - G K^T = P - [K G] [PH^T K]^T: done by the DMMA update kernel
  double2* Ub = reinterpret_cast<double2*>(a.U + (b * n + j) * (long long)(2 * DL));
  double2* Vb = reinterpret_cast<double2*>(a.V + (b * n + j) * (long long)(2 * DL));
#pragma unroll
  for (int l = 0; l < DL; l += 2) {
    Ub[l / 2] = make_double2(kj[l], kj[l + 1]);
    Ub[(DL + l) / 2] = make_double2(g[l], g[l + 1]);
    Vb[l / 2] = make_double2(ph[l], ph[l + 1]);
    Vb[(DL + l) / 2] = make_double2(kj[l], kj[l + 1]);
  }
}

__global__ void dense_init_kernel(long long B, int n, const double* P0_shared, double* P, double* nll) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nn = (long long)n * n;
  if (idx < B * nn) P[idx] = P0_shared[idx % nn];
  if (idx < B) nll[idx] = 0.0;
}

template <class Tab>
static int dense_run_tab(const odeu_plan& plan, const odeu_dense_io& io, cudaStream_t st) {
  const int D = plan.desc.ode_variant, n = 2 * D;
  const long long B = io.B, nn = (long long)n * n;
  double* J = (double*)io.workspace;
  double* M = J + B * nn;
  double* qd = M + B * nn;
  double* gq = qd + B * n;          // [n]
  double* P0d = gq + n;             // [n][n] staging of the shared initial covariance
  double* Uw = P0d + nn;            // [B][n][2 DL]
  double* Vw = Uw + B * n * 2 * DL;
  double* Sww = Vw + B * n * 2 * DL;   // [B][SW]
  // ---- fold the small host matrices
  std::vector<double> gqh(n, 0.0), R(DL * DL, 0.0);
  bool qany = false;
  if (io.Q_sqrt_diag) {
    for (int i = 0; i < n; ++i) {
      qany = qany || (io.Q_sqrt_diag[i] >= 1e-16);
      const double v = io.gamma_sqrt * io.Q_sqrt_diag[i];
      gqh[i] = v * v;
    }
  }
  int mode;
  if (plan.desc.disable_cov_update) mode = qany ? DN_Q_ONLY : DN_NONE;
  else if (qany) mode = DN_EPS_PLUS_Q;
  else mode = plan.desc.cov_fn_id == ODEU_COV_STATIC_DIAGONAL ? DN_COVFN_STATIC : DN_COVFN_DIAG;
  cudaError_t e = cudaMemcpyAsync(gq, gqh.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) { set_error("odeu_ekf_dense_run: %s", cudaGetErrorString(e)); return (int)e; }
  if (io.x != io.x0) {
    e = cudaMemcpyAsync(io.x, io.x0, sizeof(double) * B * n, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) { set_error("odeu_ekf_dense_run: %s", cudaGetErrorString(e)); return (int)e; }
  }
  if (io.P0_sqrt) {   // shared factor (host) -> P0 = P0s P0s^T broadcast to every trajectory
    std::vector<double> P0((size_t)nn, 0.0);
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < n; ++k) {
        const double aik = io.P0_sqrt[(size_t)i * n + k];
        if (aik == 0.0) continue;
        for (int j2 = 0; j2 < n; ++j2) P0[(size_t)i * n + j2] += aik * io.P0_sqrt[(size_t)j2 * n + k];
      }
    e = cudaMemcpyAsync(P0d, P0.data(), sizeof(double) * nn, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { set_error("odeu_ekf_dense_run: %s", cudaGetErrorString(e)); return (int)e; }
    cudaStreamSynchronize(st);    // P0 / gqh are host temporaries
    const long long tot = B * nn;
    dense_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(B, n, P0d, io.P, io.nll);
    count_launch();
  } else {
    cudaStreamSynchronize(st);
    e = cudaMemsetAsync(io.nll, 0, sizeof(double) * B, st);
    if (e != cudaSuccess) { set_error("odeu_ekf_dense_run: %s", cudaGetErrorString(e)); return (int)e; }
  }
  DenseCorrectArgs ca;
  ca.n = n; ca.L = io.L; ca.B = B;
  for (int l = 0; l < DL; ++l) ca.hidx[l] = l < io.L ? io.obs_index[l] : 0;
  for (int l = 0; l < io.L; ++l)
    for (int m = 0; m < io.L; ++m) {
      double s = 0.0;
      for (int k = 0; k < io.L; ++k) s += io.R_sqrt[l * io.L + k] * io.R_sqrt[m * io.L + k];
      ca.R[l * io.L + m] = s;
    }
  ca.ys = io.ys; ca.ys_per_traj = io.ys_per_trajectory; ca.flags = io.correct_flags;
  ca.ymap = (const long long*)io.xy_index_map; ca.x = io.x; ca.P = io.P; ca.nll = io.nll;
  ca.U = Uw; ca.V = Vw;

  DenseStepArgs sa;
  sa.D = D; sa.n = n; sa.B = B; sa.h = plan.desc.step_size;
  const double* th = io.theta_shared ? io.theta_shared : plan.theta_default.data();
  sa.th0 = th[0]; sa.th1 = th[1]; sa.th2 = th[2];
  sa.noise_mode = mode; sa.cov_scale = plan.desc.cov_scale; sa.gq_diag = gq;
  sa.x = io.x; sa.eps = io.eps; sa.J = J; sa.qd = qd;

  const size_t jac_smem = sizeof(double) * (n + (size_t)Tab::S * D);
  // measured on B200 (tools/bench_c5.py): 128 x 64 tiles at 2 CTAs/SM 2.56 ms/step, 128 x 128 tiles at
  // 1 CTA/SM 2.74 ms/step (the second CTA hides the prologue / epilogue of the 16-chunk k loop)
  static const bool narrow = getenv("ODEU_GEMM_BN128") == nullptr;
  double t = io.t0;
  for (long long step = 0; step < io.T; ++step) {
    sa.t = t;
    lcao_jac_kernel<Tab><<<(unsigned)B, n, jac_smem, st>>>(sa);
    if (narrow) launch_gemms<64, 4, 2>(J, io.P, M, qd, n, B, st);     // 128 x 64 tiles, 2 CTAs/SM
    else launch_gemms<128, 2, 4>(J, io.P, M, qd, n, B, st);            // 128 x 128 tiles, 1 CTA/SM
    count_launch(); count_launch(); count_launch();
    if (io.L > 0) {
      ca.step = step;
      dense_innov_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(ca, Sww);
      dense_gain_kernel<<<(unsigned)B, n, 0, st>>>(ca, Sww);
      launch_update(Uw, Vw, io.P, n, B, io.correct_flags + step, st);
      count_launch(); count_launch(); count_launch();
    }
    t += sa.h;
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("odeu_ekf_dense_run: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
  if (io.tT) *io.tT = t;
  return 0;
}

int dense_run(const odeu_plan& plan, const odeu_dense_io& io, cudaStream_t st) {
  const int D = plan.desc.ode_variant, n = 2 * D;
  if (plan.desc.ode_id != ODEU_ODE_LCAO || D < 64 || (n % GT) != 0 || n > 512) {
    set_error("odeu_ekf_dense_run: the dense path serves the oscillator chain (LCAO) with n = 2 D a multiple of 128, n <= 512");
    return -2;
  }
  if (io.B <= 0 || io.T < 0 || !io.x0 || !io.x || !io.P || !io.nll) { set_error("odeu_ekf_dense_run: B, T, x0, x, P, nll are required"); return -1; }
  if (io.L < 0 || io.L > DL) { set_error("odeu_ekf_dense_run: L=%d outside [0, %d]", io.L, DL); return -1; }
  if (io.L > 0 && (!io.obs_index || !io.R_sqrt || !io.ys || !io.correct_flags || !io.xy_index_map)) {
    set_error("odeu_ekf_dense_run: L > 0 needs obs_index, R_sqrt, ys, correct_flags, xy_index_map");
    return -1;
  }
  for (int l = 0; l < io.L; ++l)
    if (io.obs_index[l] < 0 || io.obs_index[l] >= n) { set_error("odeu_ekf_dense_run: obs_index out of range"); return -1; }
  if (plan.desc.cov_fn_id == ODEU_COV_OUTER && !plan.desc.disable_cov_update) {
    set_error("odeu_ekf_dense_run: OuterCovarianceUpdate is not served by the dense path");
    return -2;
  }
  const long long need = odeu_ekf_dense_workspace_bytes(&plan, io.B);
  if (!io.workspace || io.workspace_bytes < need) {
    set_error("odeu_ekf_dense_run: workspace of %lld bytes required", need);
    return -1;
  }
  switch (plan.desc.solver_id) {
    case ODEU_SOLVER_RKF45: return dense_run_tab<TabRKF45>(plan, io, st);
    case ODEU_SOLVER_DOPRI65: return dense_run_tab<TabDopri65>(plan, io, st);
    case ODEU_SOLVER_BS32: return dense_run_tab<TabBS32>(plan, io, st);
    case ODEU_SOLVER_HEUN_EULER: return dense_run_tab<TabHeunEuler>(plan, io, st);
    default: set_error("odeu_ekf_dense_run: unknown solver"); return -2;
  }
}

}  // namespace odeu

extern "C" {
int64_t odeu_ekf_dense_workspace_bytes(const odeu_plan* plan, int64_t B) {
  if (!plan || B <= 0) return 0;
  const long long n = plan->n;
  return (int64_t)sizeof(double) * (2 * B * n * n + B * n + n + n * n + 2 * B * n * 32 + B * (2 * 16 * 16 + 16));
}
int odeu_ekf_dense_run(const odeu_plan* plan, const odeu_dense_io* io, void* cuda_stream) {
  if (!plan || !io) { odeu::set_error("odeu_ekf_dense_run: null argument"); return -1; }
  return odeu::dense_run(*plan, *io, (cudaStream_t)cuda_stream);
}
}
