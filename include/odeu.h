/* odeu.h - C ABI of the B200-native batched EKF-over-embedded-Runge-Kutta hot path.
 *
 * This is the drop-in boundary for ONE path of f-lair/ode-uncertainty: the time loop that the
 * reference runs as `lax.scan` over jitted predict/correct closures
 *     scripts/run_filter.py:166-224                 unroll()
 *     scripts/run_parameter_estimation.py:685-796   nll()
 * with the plugins it composes
 *     src/ode/*.py                         ODE right-hand sides         -> odeu_ode_id
 *     src/solvers/{rkf45,dopri65,bs32,heun_euler}.py  RK tableaux       -> odeu_solver_id
 *     src/covariance_update_functions/*.py eps -> Q mappings            -> odeu_cov_fn_id
 *     src/filters/sqrt_ekf.py:92-197,337-376  predict / correct         -> odeu_ekf_run
 *     src/filters/particle_filter.py:73-118   perturbed-solver ensemble -> odeu_pf_run
 *     src/utils.py:109-128                 negative_log_gaussian_sqrt   -> nll output
 *
 * Conventions
 *  - plain C types only; every pointer is either HOST (small, shared configuration matrices)
 *    or DEVICE (per-trajectory batches), as marked on each field.
 *  - the library never allocates or retains caller memory: outputs are written in place into
 *    caller-provided device buffers; launches are asynchronous on the given CUDA stream.
 *  - device batch layout is component-major, trajectory-minor ("[n][B]"), so every global
 *    access of a warp is coalesced.
 *  - numerical failure is NOT an error: NaN/Inf propagate per trajectory like the reference
 *    (SURVEY section 5).  Return value: 0 ok; <0 invalid argument / unsupported combination;
 *    >0 a cudaError_t.  odeu_last_error() gives the message (thread-local).
 *  - the covariance is exchanged as the full matrix P = P_sqrt P_sqrt^T and S = S_sqrt S_sqrt^T:
 *    the reference's factors are defined only up to column signs (tests/test_utils.py:31
 *    compares c @ c.T for that reason).
 */
#ifndef ODEU_H_
#define ODEU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ODEU_VERSION 1

/* src/ode/__init__.py */
typedef enum {
  ODEU_ODE_LORENZ = 0,          /* src/ode/lorenz.py:30-54          n=3  p=3  */
  ODEU_ODE_VAN_DER_POL = 1,     /* src/ode/van_der_pol.py:22-46     n=2  p=1  */
  ODEU_ODE_LOTKA_VOLTERRA = 2,  /* src/ode/lotka_volterra.py:31-54  n=2  p=4  */
  ODEU_ODE_PENDULUM = 3,        /* src/ode/pendulum.py:22-46        n=2  p=1  */
  ODEU_ODE_LCAO = 4,            /* src/ode/lcao.py:35-63            n=2D p=3 (variant = D) */
  ODEU_ODE_HODGKIN_HUXLEY = 5,  /* src/ode/hodgkin_huxley.py:61-281 variant 0 full(8) 1 reduced-1(7) 4 reduced-4(4); p=15 */
  ODEU_ODE_MULTI_HH = 6         /* src/ode/hodgkin_huxley.py:284-439 n = num_compartments * dim(variant) */
} odeu_ode_id;

/* src/solvers/__init__.py */
typedef enum {
  ODEU_SOLVER_RKF45 = 0,      /* src/solvers/rkf45.py     */
  ODEU_SOLVER_DOPRI65 = 1,    /* src/solvers/dopri65.py   */
  ODEU_SOLVER_BS32 = 2,       /* src/solvers/bs32.py      */
  ODEU_SOLVER_HEUN_EULER = 3, /* src/solvers/heun_euler.py */
  /* implicit plugins of DiffraxSolverBuilder(name=...) (src/solvers/diffrax_solver.py:16-140): one fixed
   * step per call, Newton per stage, eps = 0.  diffrax is third-party and absent here: the published
   * methods are restated (csrc/dirk.cuh), parity unpinned.  Served by the thread-per-trajectory kernels
   * (odeu_ekf_run, odeu_ekf_grad_run, odeu_pf_run). */
  ODEU_SOLVER_KVAERNO3 = 4,       /* DiffraxSolverBuilder(name="Kvaerno3")      ESDIRK 3(2), 4 stages */
  ODEU_SOLVER_IMPLICIT_EULER = 5  /* DiffraxSolverBuilder(name="ImplicitEuler") the builder's default */
} odeu_solver_id;

/* src/covariance_update_functions/__init__.py */
typedef enum {
  ODEU_COV_DIAGONAL = 0,        /* diagonal.py:11-58        P += diag((scale*eps)^2)        */
  ODEU_COV_OUTER = 1,           /* outer.py:11-62           P += (scale*eps)(scale*eps)^T   */
  ODEU_COV_STATIC_DIAGONAL = 2  /* static_diagonal.py:11-48 P += scale^2 I (use_static_cov_fn) */
} odeu_cov_fn_id;

/* Plugin selection = what jsonargparse instantiates from class_path/init_args
 * (scripts/run_filter.py:31-47). */
typedef struct {
  int32_t ode_id;              /* odeu_ode_id */
  int32_t ode_variant;         /* HH model (0/1/4), LCAO D; else 0 */
  int32_t num_compartments;    /* ODEU_ODE_MULTI_HH only; else 0 */
  int32_t solver_id;           /* odeu_solver_id */
  double step_size;            /* SolverBuilder(step_size), src/solvers/solver.py:18-26 */
  int32_t cov_fn_id;           /* odeu_cov_fn_id */
  double cov_scale;            /* DiagonalCovarianceUpdate(scale) / static scale */
  int32_t disable_cov_update;  /* SQRT_EKF(disable_cov_update), src/filters/sqrt_ekf.py:36-43 */
} odeu_plan_desc;

typedef struct odeu_plan odeu_plan; /* opaque; immutable after creation; re-entrant per stream */

int odeu_plan_create(const odeu_plan_desc* desc, odeu_plan** out);
void odeu_plan_destroy(odeu_plan* plan);
/* flattened state dimension n = N*D and number of scalar parameters of the plan's ODE */
int odeu_plan_state_dim(const odeu_plan* plan);
int odeu_plan_num_params(const odeu_plan* plan);
/* reference default parameter values, flat, in ODEBuilder.params order (HOST out[num_params]) */
int odeu_plan_default_params(const odeu_plan* plan, double* out);

/* One batched EKF run = unroll()/nll() for B independent trajectories. */
typedef struct {
  int64_t B;                 /* trajectories / parameter sets */
  int64_t T;                 /* steps, = ceil((tN - t0) / h), scripts/run_filter.py:95 */
  double t0;
  int32_t L;                 /* observation dimension, 0 = prediction only (run_filter.py:114-121) */
  /* ---- inputs */
  const double* x0;          /* DEVICE [n][B] */
  const double* P0;          /* DEVICE [n*n][B] full covariance per trajectory, or NULL */
  const double* P0_sqrt;     /* HOST   [n][n] shared factor (used when P0 == NULL) */
  const double* theta;       /* DEVICE [p][B] per-trajectory parameters, or NULL = defaults / theta_shared */
  const double* theta_shared;/* HOST   [p] or NULL = reference defaults */
  const double* Q_sqrt;      /* HOST   [n][n] or NULL (= 0), state["Q_sqrt"], sqrt_ekf.py:77 */
  double gamma_sqrt;         /* state["gamma_sqrt"], sqrt_ekf.py:78 */
  const double* H;           /* HOST   [L][n] measurement matrix */
  const double* R_sqrt;      /* HOST   [L][L] */
  const double* ys;          /* DEVICE [T_obs][L] shared, or [T_obs][L][B] when ys_per_trajectory */
  int32_t ys_per_trajectory;
  const uint8_t* correct_flags;   /* DEVICE [T]  run_filter.py:102-103 */
  const int64_t* xy_index_map;    /* DEVICE [T]  run_filter.py:104-105 */
  int64_t save_interval;     /* 0 = no trajectory output; else slots 0, s, 2s, ... (run_filter.py:219-222) */
  int32_t skip_predict;      /* 1: apply only the measurement update in each step (single-step
                                FilterCorrect of src/filters/filter.py:31-33; API-compat/tests) */
  void* workspace;           /* DEVICE scratch of odeu_ekf_workspace_bytes() bytes, or NULL.  When
                                given (and save_interval == 0, B large) the run is scheduled
                                dynamically over (trajectory block, time segment) work items */
  int64_t workspace_bytes;
  /* ---- outputs (DEVICE, any may be NULL) */
  double* xT;                /* [n][B]    final mean */
  double* epsT;              /* [n][B]    last local error estimate */
  double* PT;                /* [n*n][B]  final covariance */
  double* yhatT;             /* [L][B]    last pre-update H x */
  double* ST;                /* [L*L][B]  last innovation covariance */
  double* nll;               /* [B]       sum over observation steps of negative_log_gaussian_sqrt */
  double* tT;                /* [1]       final time (accumulated t + h) */
  double* out_t;             /* [T_save]            T_save = T / save_interval + 1 */
  double* out_x;             /* [T_save][n][B] */
  double* out_eps;           /* [T_save][n][B] */
  double* out_P;             /* [T_save][n*n][B] */
  double* out_yhat;          /* [T_save][L][B] */
  double* out_S;             /* [T_save][L*L][B] */
  /* ---- calibration sweep (scripts/run_calibration_conrad_baseline_calibration.py:126-160):
   * one trajectory per static noise level */
  const double* cov_scale_batch;  /* DEVICE [B] per-trajectory scale of the covariance-update function
                                     (overrides the plan's cov_scale), or NULL */
  int32_t nll_nan_to_num;    /* 1: every per-step term passes through nan_to_num before it is added
                                (NaN -> 0, +/-inf -> +/-DBL_MAX), like `jnp.nan_to_num(nlls)` (:218) */
  /* ---- parameter_sensitivity (scripts/run_parameter_estimation.py:750-769): Q_sqrt = diag(w) per
   * parameter set, w from odeu_param_sensitivity.  Replaces Q_sqrt; honoured by odeu_ekf_grad_run
   * (and the NLL-only row kernels of the Hodgkin-Huxley family), rejected by the other kernels. */
  const double* Q_sqrt_diag_batch;  /* DEVICE [n][B] or NULL */
  /* ---- zero-gain guard of the measurement update (src/filters/sqrt_ekf.py:350-353).
   * ODEU_GUARD_INTENDED: K = 0 iff all(|S_sqrt| < 1e-16), evaluated on chol(S) by the full-covariance
   *   kernels (every system size).
   * ODEU_GUARD_REFERENCE: the predicate exactly as the reference writes it, all(S_sqrt < 1e-16), which
   *   is also true for a healthy factor whose entries are all negative; S_sqrt carries LAPACK's
   *   Householder signs (dlarfg), so the run is carried in FACTOR form through the same three QRs per
   *   step as the reference (sqrt_L_sum_qr, src/utils.py:233-274).  Served for n <= 4 by odeu_ekf_run;
   *   other kernels reject it.  Results are then the reference's also on inputs where the sign quirk
   *   drops observations. */
  int32_t guard_mode;        /* odeu_guard_mode */
  const double* P0_sqrt_batch; /* DEVICE [n*n][B] per-trajectory FACTOR (resume in reference mode; replaces P0) */
  double* PT_sqrt;           /* DEVICE [n*n][B] final factor with the reference's signs (reference mode), or NULL */
  double* out_P_sqrt;        /* DEVICE [T_save][n*n][B] the factor at every saved slot = state["P_sqrt"] of the
                                reference's traj_states, signs included (reference mode), or NULL */
  int64_t* guard_counts;     /* DEVICE [2][B] (reference mode) or NULL: measurement updates on which the guard
                                fired (K = 0); updates on which the verbatim and the intended predicate differ */
} odeu_ekf_io;

typedef enum { ODEU_GUARD_INTENDED = 0, ODEU_GUARD_REFERENCE = 1,
               ODEU_GUARD_INTENDED_FACTOR = 2 /* factor-form arithmetic, intended predicate (diagnostics) */
} odeu_guard_mode;

/* Scratch size for the dynamically scheduled variant of odeu_ekf_run (0 if it does not apply). */
int64_t odeu_ekf_workspace_bytes(const odeu_plan* plan, int64_t B, int64_t T);

/* Replaces unroll() (scripts/run_filter.py:166-224) and the scan of nll()
 * (scripts/run_parameter_estimation.py:771-794).  `cuda_stream` is a cudaStream_t. */
int odeu_ekf_run(const odeu_plan* plan, const odeu_ekf_io* io, void* cuda_stream);

/* EXTENSION (no reference counterpart): bootstrap-weight update for the particle
 * ensemble, logw_m += log N(y; H x_m, R).  x DEVICE [n][M]; y HOST [L]; H HOST [L][n]; R HOST
 * [L][L] (full covariance); logw DEVICE [M] in/out.  n, L <= 16. */
int odeu_pf_weight_update(int64_t M, int32_t n, int32_t L, const double* x_dev, const double* y_host,
                          const double* H_host, const double* R_host, double* logw_dev,
                          void* cuda_stream);

/* EXTENSION (bootstrap filter, BASELINE config 4; no reference counterpart): the global steps without a
 * host round trip.  Per observation, on one stream:
 *   odeu_pf_weight_reduce  logw += log N(y; H x, R) fused with this rank's triple
 *                          (max logw, sum exp(logw - max), sum exp(2 (logw - max))) -> triple_out [3]
 *   (caller: one all-gather of the G triples - NCCL - on the same stream)
 *   odeu_pf_normalize      lse over the G triples; logw -= lse; stats = {lse, ESS, resample flag, running
 *                          log-likelihood (accumulated)} as DEVICE scalars; pack [M][n+1] rows (x, w)
 *   (caller: one all-gather of the packed rows and an inclusive cumulative sum of the weights)
 *   odeu_pf_resample       systematic resampling (slot j of the GLOBAL ensemble takes the particle whose
 *                          CDF interval contains (j + u0) / M_total), predicated on stats[2]; writes the
 *                          [n][M] layout odeu_pf_run reads and resets logw to -log M_total
 * scratch: odeu_pf_reduce_scratch_bytes(M) bytes, zeroed once by the caller. */
int64_t odeu_pf_reduce_scratch_bytes(int64_t M);
int odeu_pf_weight_reduce(int64_t M, int32_t n, int32_t L, const double* x_dev, const double* y_host,
                          const double* H_host, const double* R_host, double* logw_dev,
                          double* triple_out_dev, void* scratch_dev, void* cuda_stream);
int odeu_pf_normalize(int64_t M, int64_t M_total, int32_t n, int32_t G, const double* triples_dev,
                      const double* x_dev, double* logw_dev, double* pack_dev, double* stats_dev,
                      double* ess_hist_dev, double* flag_hist_dev, double ess_frac, void* cuda_stream);
int odeu_pf_resample(int64_t M, int64_t M_total, int64_t slot_lo, int32_t n, double u0,
                     const double* stats_dev, const double* cdf_dev, const double* pack_dev,
                     double* x_new_dev, double* logw_dev, void* cuda_stream);
/* The last two steps in one call: inclusive cumulative sum of the gathered weights (read straight off the packed
 * rows) into scan scratch (odeu_pf_scan_bytes(M_total) bytes) + odeu_pf_resample; when the device-side decision
 * is "keep", x_old [n][M] is carried over into x_new, so the caller alternates two ensemble buffers. */
int64_t odeu_pf_scan_bytes(int64_t M_total);
int odeu_pf_scan_resample(int64_t M, int64_t M_total, int64_t slot_lo, int32_t n, double u0,
                          const double* stats_dev, const double* pack_dev, const double* x_old_dev,
                          double* x_new_dev, double* logw_dev, void* scan_dev, int64_t scan_bytes,
                          void* cuda_stream);

/* Peer-memory variant of the global steps (one process per GPU; every rank's buffers are mapped into every other
 * rank over NVLink / NVSwitch, e.g. by torch.distributed._symmetric_memory or CUDA IPC): no collective library on
 * the data path.  Pointer tables are HOST arrays of G device pointers (entry q = rank q's buffer), G <= 16.
 *   odeu_pf_publish_triple      stores this rank's triple into slot `rank` of every rank's triples [G][3]
 *   (caller: cross-rank barrier)
 *   odeu_pf_normalize_w         odeu_pf_normalize + the weights once more as a contiguous array w [M]; pack and w
 *                               are this rank's SYMMETRIC buffers
 *   (caller: cross-rank barrier)
 *   odeu_pf_scan_resample_peer  CDF scan streaming the G weight arrays from their owners (8 B per particle), then
 *                               systematic resampling that fetches only the ancestor rows it needs from the owning
 *                               rank (one 32-byte sector each at n = 3) - an all-gather would deliver the whole
 *                               ensemble to every rank.  M_total = G * M. */
int odeu_pf_publish_triple(const double* triple_dev, double* const* peer_triples_host, int32_t rank, int32_t G,
                           void* cuda_stream);
int odeu_pf_normalize_w(int64_t M, int64_t M_total, int32_t n, int32_t G, const double* triples_dev,
                        const double* x_dev, double* logw_dev, double* pack_dev, double* w_dev,
                        double* stats_dev, double* ess_hist_dev, double* flag_hist_dev, double ess_frac,
                        void* cuda_stream);
int odeu_pf_scan_resample_peer(int64_t M, int64_t M_total, int64_t slot_lo, int32_t n, int32_t G, double u0,
                               const double* stats_dev, const double* const* peer_pack_host,
                               const double* const* peer_w_host, const double* x_old_dev, double* x_new_dev,
                               double* logw_dev, void* scan_dev, int64_t scan_bytes, void* cuda_stream);

/* NLL and its parameter gradient for B parameter sets: replaces jax.value_and_grad(nll) as the
 * optimiser calls it (scripts/run_parameter_estimation.py:599, nll :685-796).  Forward-mode
 * tangents over the requested parameters (at most 32), fused with the filter loop.  Uses from
 * `io`: B, T, t0, L, x0, P0_sqrt, theta / theta_shared, Q_sqrt, gamma_sqrt, H, R_sqrt, ys,
 * ys_per_trajectory, correct_flags, xy_index_map; writes io->nll [B], io->xT [n][B] (optional)
 * and grad->grad.  The derivative is w.r.t. the PHYSICAL parameter theta_j; the caller applies
 * d theta / d theta_norm = (max - min) (src/utils.py:156-178). */
typedef struct {
  int32_t p_opt;             /* number of differentiated parameters, 1..32 */
  const int32_t* idx;        /* HOST [p_opt] flat parameter indices (ODEBuilder.params order) */
  const double* x0_tangent;  /* DEVICE [p_opt][n][B] d x0 / d theta_j (initial_state_parametrized,
                                run_parameter_estimation.py:744-748) or NULL = 0 */
  double* grad;              /* DEVICE [p_opt][B] d NLL / d theta_j */
  const double* Q_sqrt_diag_tangent; /* DEVICE [p_opt][n][B] d w / d theta_j for io->Q_sqrt_diag_batch
                                (w_tangent of odeu_param_sensitivity), or NULL = 0 */
} odeu_grad_io;

int odeu_ekf_grad_run(const odeu_plan* plan, const odeu_ekf_io* io, const odeu_grad_io* grad,
                      void* cuda_stream);

/* Lock-step projected L-BFGS over R restarts on the device: the optimiser loop of
 * scripts/run_parameter_estimation.py:599-667 (SciPy L-BFGS-B through jaxopt, one process per restart)
 * without leaving the GPU; src/utils.py:15-36 is the reference's own projected L-BFGS.  Call pattern:
 *   zt = starting points; evaluate (odeu_ekf_grad_run) -> odeu_lbfgs_step(first = 1)
 *   repeat: evaluate the trial points zt -> odeu_lbfgs_step(first = 0)
 * Every call advances every restart by one evaluation (Armijo backtracking on the projected path, m = 10
 * curvature pairs, SciPy's stopping rules pgtol / ftol / maxiter).  workspace (zeroed before the first
 * call) layout in doubles: z [R][p], g [R][p], d [R][p], f [R], alpha [R], S [R][10][p], Y [R][10][p],
 * rho [R][10], then int32 meta [R][6] = {pairs, head, iterations, evaluations, status, ls trials};
 * status: 0 running, 1 |proj grad| <= pgtol, 2 relative decrease <= ftol, 3 maxiter, 4 line search, 5 NaN. */
int64_t odeu_lbfgs_workspace_doubles(int32_t R, int32_t p);
int odeu_lbfgs_step(int32_t R, int32_t p, int32_t maxiter, int32_t first, double pgtol, double ftol,
                    double* workspace_dev, double* zt_dev, const double* ft_dev, const double* gt_dev,
                    const double* scale_dev, void* cuda_stream);

/* Parameter-sensitivity weights of the process noise, replaces the `if parameter_sensitivity:` block
 * of nll() (scripts/run_parameter_estimation.py:750-769): one solver step from (t0, x0),
 *   w_i = sum_{k in idx} |d x1_i / d theta_k|,   w <- sqrt(n) w / |w|_2,   Q_sqrt = diag(w).
 * The reference evaluates it inside the differentiated loss; w_tangent carries d w / d theta_j
 * (second derivatives of the step; with x0_tangent also the dependence through
 * initial_state_parametrized) for odeu_grad_io.Q_sqrt_diag_tangent.  Derivatives are w.r.t. the
 * PHYSICAL parameters, like odeu_ekf_grad_run. */
typedef struct {
  int64_t B;
  double t0;
  const double* x0;          /* DEVICE [n][B] */
  const double* theta;       /* DEVICE [p][B] or NULL = theta_shared / defaults */
  const double* theta_shared;/* HOST   [p] or NULL = reference defaults */
  int32_t p_opt;             /* number of optimised parameters, 1..32 */
  const int32_t* idx;        /* HOST [p_opt] flat parameter indices */
  const double* x0_tangent;  /* DEVICE [p_opt][n][B] or NULL */
  double* w;                 /* DEVICE [n][B] out */
  double* w_tangent;         /* DEVICE [p_opt][n][B] out, or NULL */
} odeu_sens_io;

int odeu_param_sensitivity(const odeu_plan* plan, const odeu_sens_io* io, void* cuda_stream);

/* Large-state EKF (BASELINE config 5): the oscillator chain of src/ode/lcao.py:51-61 generalised to
 * D oscillators (plan: ode_id = ODEU_ODE_LCAO, ode_variant = D >= 64, n = 2 D a multiple of 128),
 * where P <- J P J^T + Q is a dense contraction executed with FP64 tensor-core MMAs.  Replaces the
 * same `unroll()` / `nll()` loop as odeu_ekf_run (scripts/run_filter.py:166-224) for that shape.
 * Layout differs from odeu_ekf_io: matrices are stored PER TRAJECTORY, row-major.
 * The measurement matrix must select state components (rows of the identity: every
 * measurement_matrix the reference ships has this form); R may be any L x L factor. */
typedef struct {
  int64_t B, T;
  double t0;
  int32_t L;                      /* number of observed components, 0..16 */
  const double* x0;               /* DEVICE [B][n] */
  double* x;                      /* DEVICE [B][n] state, updated in place (may alias x0) */
  double* P;                      /* DEVICE [B][n][n] covariance, in/out (symmetric) */
  const double* P0_sqrt;          /* HOST [n][n] shared initial factor, or NULL: P already holds P0 */
  const double* theta_shared;     /* HOST [3] (lin, cubic, coupling) or NULL = reference defaults */
  const double* Q_sqrt_diag;      /* HOST [n] diagonal of Q_sqrt (tempering), or NULL */
  double gamma_sqrt;
  const int32_t* obs_index;       /* HOST [L] observed state components */
  const double* R_sqrt;           /* HOST [L][L] */
  const double* ys;               /* DEVICE [T_obs][L] shared or [T_obs][B][L] */
  int32_t ys_per_trajectory;
  const uint8_t* correct_flags;   /* DEVICE [T] */
  const int64_t* xy_index_map;    /* DEVICE [T] */
  double* eps;                    /* DEVICE [B][n] last embedded error estimate, or NULL */
  double* nll;                    /* DEVICE [B] negative log-likelihood */
  double* tT;                     /* HOST final time, or NULL */
  void* workspace;                /* DEVICE, odeu_ekf_dense_workspace_bytes(plan, B) bytes */
  int64_t workspace_bytes;
} odeu_dense_io;

int64_t odeu_ekf_dense_workspace_bytes(const odeu_plan* plan, int64_t B);
int odeu_ekf_dense_run(const odeu_plan* plan, const odeu_dense_io* io, void* cuda_stream);


/* Perturbed-solver particle ensemble (src/filters/particle_filter.py:24-118; Conrad et al.
 * baseline): M independent RK steps, after each step x += p with p ~ N(0, covfn(0, eps_m));
 * the particle with GLOBAL index 0 is noise-free (:104-105).  The reference draws from JAX's
 * threefry stream (sample-level parity impossible); here the normals come from
 * Philox4x32-10 keyed by `seed` with counter (global particle index, global step index), so
 * results do not depend on how particles are sharded over GPUs or on chunked (resumed) runs. */
typedef struct {
  int64_t M;                 /* particles on this rank */
  int64_t T;                 /* steps */
  double t0;
  const double* x0;          /* DEVICE [n][M], or NULL -> x0_shared */
  const double* x0_shared;   /* HOST   [n] (ParticleFilter.init_state broadcasts x0, :52-63) */
  const double* theta_shared;/* HOST   [p] or NULL = reference defaults */
  uint64_t seed;
  int64_t particle_offset;   /* global index of local particle 0 (sharding) */
  int64_t step_offset;       /* global index of local step 0 (resume) */
  int64_t save_interval;     /* 0 = none */
  int32_t noise_free;        /* 1: no perturbation at all = plain RK solver (src/solvers/rksolver.py:113-155) */
  double* xT;                /* DEVICE [n][M] */
  double* epsT;              /* DEVICE [n][M] */
  double* tT;                /* DEVICE [1] */
  double* out_t;             /* DEVICE [T_save] */
  double* out_x;             /* DEVICE [T_save][n][M] */
  double* out_eps;           /* DEVICE [T_save][n][M] */
} odeu_pf_io;

/* Replaces unroll() driven by ParticleFilter.build_predict (particle_filter.py:73-118). */
int odeu_pf_run(const odeu_plan* plan, const odeu_pf_io* io, void* cuda_stream);

/* ODE right-hand side dx/dt = f(t, x, theta) for a batch (the `ODE` callable of
 * src/ode/ode.py:6-7).  x, dx: DEVICE [n][B]; theta: DEVICE [p][B] or NULL -> theta_shared (HOST
 * [p]) or NULL -> reference defaults. */
int odeu_ode_rhs(const odeu_plan* plan, int64_t B, double t, const double* x, const double* theta,
                 const double* theta_shared, double* dx, void* cuda_stream);

/* FP64-pipe micro-benchmark (roofline denominator for the EKF kernels): launches `blocks` x
 * `threads` threads each running 8 independent chains of `iters` dependent DFMAs; writes the
 * number of flops issued to *flops_out (HOST).  Time it with CUDA events on `cuda_stream`. */
int odeu_bench_dfma(int64_t iters, int32_t blocks, int32_t threads, double* scratch_dev,
                    double* flops_out, void* cuda_stream);

/* number of kernel launches issued by this library in this process (bench bookkeeping) */
int64_t odeu_launch_count(void);

/* last error message of the calling thread; returns the message length */
size_t odeu_last_error(char* buf, size_t buflen);

int odeu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ODEU_H_ */
