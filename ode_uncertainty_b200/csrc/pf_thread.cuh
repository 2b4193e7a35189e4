// Perturbed-solver particle ensemble, one thread per particle
// (reference: src/filters/particle_filter.py:73-118).
//
// Per step: plain RK step, then x += p, p ~ N(0, covfn(0, eps)) with
//   Diagonal : cov = diag((scale eps)^2)        -> p_i = scale eps_i z_i      (diagonal.py:24-41)
//   Outer    : cov = (scale eps)(scale eps)^T   -> p   = scale eps z          (outer.py:24-42)
//   Static   : cov = scale^2 I                  -> p_i = scale z_i            (static_diagonal.py:14-29)
// (the reference factorises the dense [M,n,n] covariance by SVD, particle_filter.py:93-102; for
// these three plugins the factor is known in closed form).  Global particle 0 stays noise-free.
// Normals: Philox4x32-10, key = seed, counter = (global particle, global step, draw, 0),
// Box-Muller with single-precision transcendentals (see Philox::normal2).
#pragma once
#include "ekf_core.cuh"
#include "dirk.cuh"

namespace odeu {

struct Philox {
  static constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __host__ __device__ static inline void round(unsigned (&c)[4], unsigned k0, unsigned k1) {
    const unsigned long long p0 = (unsigned long long)M0 * c[0];
    const unsigned long long p1 = (unsigned long long)M1 * c[2];
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c[1] ^ k0;
    const unsigned n1 = (unsigned)p1;
    const unsigned n2 = (unsigned)(p0 >> 32) ^ c[3] ^ k1;
    const unsigned n3 = (unsigned)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __host__ __device__ static inline void gen(unsigned long long seed, unsigned long long particle,
                                             unsigned long long step, unsigned draw,
                                             unsigned (&out)[4]) {
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
    // counter: particle (40 bits) | step (48 bits) | draw (8 bits) spread over 128 bits
    unsigned c[4] = {(unsigned)particle, (unsigned)(particle >> 32) | (draw << 24),
                     (unsigned)step, (unsigned)(step >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(c, k0, k1);
      k0 += W0; k1 += W1;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
  // Two standard normals from one 128-bit block: Box-Muller with SINGLE-PRECISION transcendentals.
  //   radius: E = -log(u), u = (k + 1/2) 2^-23 from 23 bits (exact in float, never 0 or 1); k = 0
  //           (probability 2^-23) extends the tail with 32 more bits, so radii reach 8.7 sigma like
  //           a 55-bit uniform;  angle: 32 bits, theta in [-pi, pi).
  // Why float: the draws only have to be standard normal to statistical resolution (their effect is
  // x += scale * eps * z with eps ~ 1e-9 |x|); in double the generator cost ~45 FP64-pipe
  // instructions per normal against ~110 for the whole RK step and capped the ensemble kernel at
  // 26 % of the FP64 peak.  logf / sqrtf / sincosf run on the FP32 and MUFU pipes, which
  // the RK step leaves idle.  The resulting z carries ~2^-22 relative error, far below any test or
  // Monte-Carlo resolution; sharding / resume invariance is untouched (same counter keying).
  __host__ __device__ static inline void normal2(unsigned long long seed, unsigned long long particle,
                                                 unsigned long long step, unsigned draw,
                                                 double* z0, double* z1) {
    unsigned r[4];
    gen(seed, particle, step, draw, r);
    const unsigned k = r[0] >> 9;
    float E;
    if (k != 0) {
      const float u = ((float)k + 0.5f) * 1.1920928955078125e-07f;          // 2^-23
#ifdef __CUDA_ARCH__
      E = -__logf(u);
#else
      E = -logf(u);
#endif
    } else {              // u < 2^-23: -log(2^-23 v), v = (r1 + 1/2) 2^-32
      const float v = ((float)(r[1] >> 8) + 0.5f) * 5.9604644775390625e-08f;  // 24 bits, exact
      E = 15.942385152878742f - logf(v);
    }
    const float rad = sqrtf(2.0f * E);
    const float th = ((float)(r[2] >> 8) - 8388608.0f) * 3.7450702829239286e-07f;   // (k - 2^23) 2 pi / 2^24
    float sn, cs;
#ifdef __CUDA_ARCH__
    __sincosf(th, &sn, &cs);
#else
    sn = sinf(th); cs = cosf(th);
#endif
    *z0 = (double)(rad * cs);
    *z1 = (double)(rad * sn);
  }
};

template <int NX, int NP>
struct PfArgs {
  long long M, T;
  double t0, h;
  int cov_fn;
  double cov_scale;
  unsigned long long seed;
  long long particle_offset, step_offset, save_interval;
  int noise_free;
  const double* x0;
  double* xT; double* epsT; double* tT; double* out_t; double* out_x; double* out_eps;
  double x0s[NX];
  double theta_shared[NP];
  ScaledTableau st;        // tableau coefficients as constant-bank operands (explicit solvers)
};

template <class Ode, class Tab>
ODEU_HD void pf_particle(const PfArgs<Ode::NX, Ode::NP>& a, const long long m) {
  constexpr int n = Ode::NX;
  constexpr int NP = Ode::NP;
  const long long M = a.M;
  const unsigned long long gid = (unsigned long long)(a.particle_offset + m);
  double x[n], eps[n], th[NP];
#pragma unroll
  for (int i = 0; i < n; ++i) { x[i] = a.x0 ? a.x0[i * M + m] : a.x0s[i]; eps[i] = 0.0; }
#pragma unroll
  for (int k = 0; k < NP; ++k) th[k] = a.theta_shared[k];
  double t = a.t0;
  const long long si = a.save_interval;
  if (si > 0) {
#pragma unroll
    for (int i = 0; i < n; ++i) {
      if (a.out_x) a.out_x[i * M + m] = x[i];
      if (a.out_eps) a.out_eps[i * M + m] = 0.0;
    }
    if (m == 0 && a.out_t) a.out_t[0] = t;
  }
  long long next_save = si, slot = 1;
  double spare = 0.0;          // second normal of the last Box-Muller block of an even step
  bool have_spare = false;
  for (long long step = 0; step < a.T; ++step) {
    double xn[n];
    if constexpr (is_implicit<Tab>::value) {
      double Jd[n][n];      // (the step Jacobian comes with the Newton solve; unused here)
      dirk_step_generic<Ode, Tab, double>(t, a.h, x, th, xn, eps, Jd);
    } else {
      rk_step_plain_st<Ode, Tab>(t, a.h, a.st, x, th, xn, eps);
    }
    t = t + a.h;
    if (gid != 0 && !a.noise_free) {
      const unsigned long long gstep = (unsigned long long)(a.step_offset + step);
      // NZ standard normals per step (one for the rank-one Outer plugin, n otherwise).  Box-Muller
      // yields them in pairs: an odd NZ would waste one per step, so two consecutive steps
      // (2q, 2q + 1) share the NZ blocks keyed (particle, q): the even step takes normals 0..NZ-1
      // and keeps normal NZ for the odd step (recomputed if a run starts on an odd step), which
      // adds normals NZ+1..2NZ-1.  Keyed by GLOBAL particle and step: sharding/resume invariant.
      const bool outer = a.cov_fn == COV_OUTER;
      double z[n + 1];
      const int nz = outer ? 1 : n;
      if (nz & 1) {
        const unsigned long long q = gstep >> 1;
        const int nb = (nz + 1) / 2;                    // blocks of the even step
        if (!(gstep & 1)) {
#pragma unroll
          for (int i = 0; i < (n + 1) / 2; ++i)
            if (i < nb) Philox::normal2(a.seed, gid, q, (unsigned)i, &z[2 * i], &z[2 * i + 1]);
          spare = z[nz];
          have_spare = true;
        } else {
          if (!have_spare) {
            double dummy;
            Philox::normal2(a.seed, gid, q, (unsigned)(nb - 1), &dummy, &spare);
          }
          z[0] = spare;
          have_spare = false;
#pragma unroll
          for (int i = 0; i < n / 2; ++i)
            if (i < nz / 2) Philox::normal2(a.seed, gid, q, (unsigned)(nb + i), &z[1 + 2 * i], &z[2 + 2 * i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < n / 2; ++i) Philox::normal2(a.seed, gid, gstep, (unsigned)i, &z[2 * i], &z[2 * i + 1]);
      }
      if (outer) {
#pragma unroll
        for (int i = 0; i < n; ++i) xn[i] = fma(a.cov_scale * eps[i], z[0], xn[i]);
      } else {
#pragma unroll
        for (int i = 0; i < n; ++i) {
          const double s0 = (a.cov_fn == COV_DIAGONAL) ? a.cov_scale * eps[i] : a.cov_scale;
          xn[i] = fma(s0, z[i], xn[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < n; ++i) x[i] = xn[i];
    if (si > 0 && step + 1 == next_save) {
#pragma unroll
      for (int i = 0; i < n; ++i) {
        if (a.out_x) a.out_x[(slot * n + i) * M + m] = x[i];
        if (a.out_eps) a.out_eps[(slot * n + i) * M + m] = eps[i];
      }
      if (m == 0 && a.out_t) a.out_t[slot] = t;
      ++slot;
      next_save += si;
    }
  }
#pragma unroll
  for (int i = 0; i < n; ++i) {
    if (a.xT) a.xT[i * M + m] = x[i];
    if (a.epsT) a.epsT[i * M + m] = eps[i];
  }
  if (m == 0 && a.tT) a.tT[0] = t;
}

template <class Ode, class Tab, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
pf_thread_kernel(const __grid_constant__ PfArgs<Ode::NX, Ode::NP> a) {
  const long long m = (long long)blockIdx.x * BLOCK + threadIdx.x;
  if (m >= a.M) return;
  pf_particle<Ode, Tab>(a, m);
}

}  // namespace odeu

namespace odeu {
// dx/dt = f(t, x, theta) for a batch: the `ODE` callable of src/ode/ode.py:6-7.
template <int NP>
struct RhsArgs {
  long long B;
  double t;
  const double* x; const double* theta; double* dx;
  double theta_shared[NP];
};
template <class Ode>
__global__ void __launch_bounds__(128) ode_rhs_kernel(const __grid_constant__ RhsArgs<Ode::NP> a) {
  const long long b = (long long)blockIdx.x * 128 + threadIdx.x;
  if (b >= a.B) return;
  double x[Ode::NX], dx[Ode::NX], th[Ode::NP];
#pragma unroll
  for (int i = 0; i < Ode::NX; ++i) x[i] = a.x[i * a.B + b];
#pragma unroll
  for (int k = 0; k < Ode::NP; ++k) th[k] = a.theta ? a.theta[k * a.B + b] : a.theta_shared[k];
  Ode::rhs(a.t, x, th, dx);
#pragma unroll
  for (int i = 0; i < Ode::NX; ++i) a.dx[i * a.B + b] = dx[i];
}
}  // namespace odeu
