// TEST INFRASTRUCTURE ONLY (never linked into libodeu.so, never on the product path).
//
// Compiles the *same* per-trajectory source the CUDA kernels run (ekf_trajectory / pf_particle,
// __host__ __device__) for the CPU, so the `-m "not gpu"` suite can check the kernel arithmetic
// against the oracle in a container without a GPU.  All pointers of the io structs are HOST
// pointers here.  The shipped library has no such entry point.
#include "../ode_uncertainty_b200/csrc/launch.cuh"
#include "../ode_uncertainty_b200/csrc/launch_grad.cuh"

#include <vector>

namespace odeu {
static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  g_err = buf;
}
void count_launch() {}

// Sequential replay of one launch of ekf_rows_kernel (same barrier intervals, same order).
template <class Ode, class Tab, class S>
int emu_rows(GradArgs<Ode::NX, Ode::NP>& a, double* PT) {
  if constexpr (has_rows<Ode>::value) {
    constexpr int n = Ode::NX;
    constexpr int TB = 4;                     // small CTA: ragged tail exercised
    fill_rt_tableau<Tab>(a);
    fill_rows_schedule<Tab>(a);
    using Th = RowThread<Ode, Tab, S, TB, 0>;
    using SM = RowsSmem<Ode, Tab, S, TB>;
    constexpr int UPC = TB * lanes_of<S>::value;       // units per CTA
    const long long units = a.B * (a.p_opt > 0 ? a.p_opt : 1);
    std::vector<S> sm((size_t)SM::total);
    std::vector<Th> th((size_t)n * TB);
    for (long long cta = 0; cta < (units + UPC - 1) / UPC; ++cta) {
      for (int k = 0; k < n * TB; ++k) {
        const int tl = k % TB, g = (k / TB) % Ode::ROW_GROUPS, q = k / (TB * Ode::ROW_GROUPS);
        th[k].init(a, cta * UPC + tl, tl, g, q, sm.data());
      }
      for (auto& t : th) t.save_initial(a, sm.data());
      for (long long step = 0; step < a.T; ++step) {
        for (int i = 0; i < a.rt_S; ++i) {
          for (auto& t : th) t.stage_a_rt(a, i, sm.data());
          for (auto& t : th) t.stage_b(a, i, sm.data());
          for (auto& t : th) t.stage_c_rt(a, i, sm.data());
        }
        for (auto& t : th) t.phase_x(a, sm.data());
        for (auto& t : th) t.phase_mp(a, sm.data());
        for (auto& t : th) t.phase_noise_outer(a, sm.data());
        if (a.has_obs && a.flags[step]) {
          const long long oi = a.ymap[step];
          for (auto& t : th) t.phase_pht(a, sm.data());
          for (auto& t : th) {
            S y[ROWS_LMAX];
            for (int l = 0; l < a.L; ++l) {
              y[l] = S(0.0);
              for (int u = 0; u < lanes_of<S>::value; ++u)
                lane_set(y[l], u, a.ys_per_traj ? a.ys[(oi * a.L + l) * a.B + t.b[u]] : a.ys[oi * a.L + l]);
            }
            t.phase_gain(a, y, sm.data());
          }
          for (auto& t : th) t.phase_update(a, sm.data());
        }
        for (auto& t : th) t.phase_store(sm.data());
        for (auto& t : th) t.save_step(a, step);
      }
      for (auto& t : th) t.finish(a, PT, sm.data());
    }
    return 0;
  } else {
    return -2;
  }
}

template <class Ode, class Tab>
int emu_ekf(const odeu_plan& plan, const odeu_ekf_io& io) {
  if (rows_eligible<Ode, Tab, double>(io) && (plan.desc.ode_id != ODEU_ODE_HODGKIN_HUXLEY || plan.desc.ode_variant != 4)) {
    GradArgs<Ode::NX, Ode::NP> g;                 // same routing as odeu_ekf_run
    if (int rc = fill_grad_args<Ode>(plan, io, nullptr, g)) return rc;
    if constexpr (std::is_same<Tab, TabRKF45>::value && rows_static_ok<Ode, Tab, V2d>()) {
      if (getenv("ODEU_ROWS_2WIDE")) return emu_rows<Ode, Tab, V2d>(g, io.PT);       // same routing as launch_rows
    }
    return emu_rows<Ode, Tab, double>(g, io.PT);
  }
  EkfArgs<Ode::NX, Ode::NP> a;
  if (int rc = fill_ekf_args<Ode>(plan, io, a)) return rc;
  fill_scaled_tableau<Tab>(plan.desc.step_size, a.st);
  constexpr int n = Ode::NX;
  constexpr int KC = LaunchCfg<Ode>::KC;
  const int lk = select_lk<Ode>(io);
  auto run = [&](long long b, const Segment& sg) {
    if constexpr (is_implicit<Tab>::value) {               // same routing as launch_ekf
      if constexpr (n <= 4) {
        if (io.guard_mode != ODEU_GUARD_INTENDED) ekf_trajectory<Ode, Tab, KC, -1, 2>(a, b, sg);
        else ekf_trajectory<Ode, Tab, KC, -1>(a, b, sg);
      } else {
        ekf_trajectory<Ode, Tab, KC, -1>(a, b, sg);
      }
    } else
    if constexpr (n <= 4) {
      if (io.guard_mode != ODEU_GUARD_INTENDED) {          // same routing as launch_ekf
        if (factor_fast_ok<Ode>(a, io, lk)) {
          if (lk == 1) ekf_trajectory<Ode, Tab, KC, 1, 1>(a, b, sg);
          else ekf_trajectory<Ode, Tab, KC, n, 1>(a, b, sg);
        } else {
          ekf_trajectory<Ode, Tab, KC, -1, 2>(a, b, sg);
        }
      } else
      if (lk == 0) ekf_trajectory<Ode, Tab, KC, 0>(a, b, sg);
      else if (lk == 1) ekf_trajectory<Ode, Tab, KC, 1>(a, b, sg);
      else if (lk == n) ekf_trajectory<Ode, Tab, KC, n>(a, b, sg);
      else ekf_trajectory<Ode, Tab, KC, -1>(a, b, sg);
    } else {
      ekf_trajectory<Ode, Tab, KC, -1>(a, b, sg);
    }
  };
  if (io.workspace && io.workspace_bytes > 0 && io.save_interval == 0) {
    // sequential replay of the dynamic scheduler's work items (segment-major), with a
    // caller-chosen segment length smuggled through workspace_bytes' low bits is NOT needed:
    // use 7 steps per segment so short test runs cross many boundaries
    const long long seg_len = 7, nseg = (io.T + seg_len - 1) / seg_len;
    for (long long seg = 0; seg < nseg; ++seg)
      for (long long b = 0; b < io.B; ++b) {
        Segment sg = {seg * seg_len, (seg + 1 == nseg) ? (long long)io.T : (seg + 1) * seg_len, seg == 0,
                      seg + 1 == nseg, (double*)io.workspace, nullptr, nullptr};
        run(b, sg);
      }
    return 0;
  }
  const Segment whole = {0, (long long)io.T, true, true, nullptr, nullptr, nullptr};
  for (long long b = 0; b < io.B; ++b) run(b, whole);
  return 0;
}
template <class Ode, class Tab>
int emu_pf(const odeu_plan& plan, const odeu_pf_io& io) {
  PfArgs<Ode::NX, Ode::NP> a;
  if (int rc = fill_pf_args<Ode>(plan, io, a)) return rc;
  if constexpr (!is_implicit<Tab>::value) fill_scaled_tableau<Tab>(plan.desc.step_size, a.st);
  for (long long m = 0; m < io.M; ++m) pf_particle<Ode, Tab>(a, m);
  return 0;
}

// host replay of launch_grad: same argument folding, then the per-(trajectory, chunk) body
template <class Ode, class Tab>
int emu_grad(const odeu_plan& plan, const odeu_ekf_io& io, const odeu_grad_io& g);

// host replay of launch_sens: the per-(parameter set, direction) body of sens.cuh
template <class Ode, class Tab>
int emu_sens(const odeu_plan& plan, const odeu_sens_io& s) {
  SensArgs<Ode::NP> a;
  a.B = s.B; a.t0 = s.t0; a.h = plan.desc.step_size; a.p_opt = s.p_opt;
  for (int j = 0; j < ODEU_MAX_GRAD; ++j) a.idx[j] = j < s.p_opt ? s.idx[j] : -1;
  a.x0 = s.x0; a.x0_tan = s.x0_tangent; a.theta = s.theta; a.w = s.w; a.w_tan = s.w_tangent;
  for (int k = 0; k < Ode::NP; ++k) a.theta_shared[k] = s.theta_shared ? s.theta_shared[k] : plan.theta_default[k];
  for (int j = 0; j < (s.w_tangent ? s.p_opt : 1); ++j)
    for (long long b = 0; b < s.B; ++b) param_sens_unit<Ode, Tab>(a, b, j);
  return 0;
}

template <class Ode>
int emu_solver(const odeu_plan& plan, const odeu_ekf_io* e, const odeu_pf_io* p, const odeu_grad_io* g = nullptr,
               const odeu_sens_io* s = nullptr) {
  if (s) {
    switch (plan.desc.solver_id) {
      case ODEU_SOLVER_RKF45: return emu_sens<Ode, TabRKF45>(plan, *s);
      case ODEU_SOLVER_DOPRI65: return emu_sens<Ode, TabDopri65>(plan, *s);
      case ODEU_SOLVER_BS32: return emu_sens<Ode, TabBS32>(plan, *s);
      case ODEU_SOLVER_HEUN_EULER: return emu_sens<Ode, TabHeunEuler>(plan, *s);
      case ODEU_SOLVER_KVAERNO3: return emu_sens<Ode, TabKvaerno3>(plan, *s);
      case ODEU_SOLVER_IMPLICIT_EULER: return emu_sens<Ode, TabImplicitEuler>(plan, *s);
    }
    return -2;
  }
  if (g) {
    switch (plan.desc.solver_id) {
      case ODEU_SOLVER_RKF45: return emu_grad<Ode, TabRKF45>(plan, *e, *g);
      case ODEU_SOLVER_DOPRI65: return emu_grad<Ode, TabDopri65>(plan, *e, *g);
      case ODEU_SOLVER_BS32: return emu_grad<Ode, TabBS32>(plan, *e, *g);
      case ODEU_SOLVER_HEUN_EULER: return emu_grad<Ode, TabHeunEuler>(plan, *e, *g);
      case ODEU_SOLVER_KVAERNO3: return emu_grad<Ode, TabKvaerno3>(plan, *e, *g);
      case ODEU_SOLVER_IMPLICIT_EULER: return emu_grad<Ode, TabImplicitEuler>(plan, *e, *g);
    }
    return -2;
  }
  switch (plan.desc.solver_id) {
    case ODEU_SOLVER_RKF45: return e ? emu_ekf<Ode, TabRKF45>(plan, *e) : emu_pf<Ode, TabRKF45>(plan, *p);
    case ODEU_SOLVER_DOPRI65: return e ? emu_ekf<Ode, TabDopri65>(plan, *e) : emu_pf<Ode, TabDopri65>(plan, *p);
    case ODEU_SOLVER_BS32: return e ? emu_ekf<Ode, TabBS32>(plan, *e) : emu_pf<Ode, TabBS32>(plan, *p);
    case ODEU_SOLVER_HEUN_EULER: return e ? emu_ekf<Ode, TabHeunEuler>(plan, *e) : emu_pf<Ode, TabHeunEuler>(plan, *p);
    case ODEU_SOLVER_KVAERNO3: return e ? emu_ekf<Ode, TabKvaerno3>(plan, *e) : emu_pf<Ode, TabKvaerno3>(plan, *p);
    case ODEU_SOLVER_IMPLICIT_EULER: return e ? emu_ekf<Ode, TabImplicitEuler>(plan, *e) : emu_pf<Ode, TabImplicitEuler>(plan, *p);
  }
  return -2;
}

static int emu_dispatch(const odeu_plan_desc& d, const double* theta_default, int p,
                        const odeu_ekf_io* e, const odeu_pf_io* pf, const odeu_grad_io* g = nullptr,
                        const odeu_sens_io* s = nullptr) {
  odeu_plan plan;
  odeu_plan_desc dm = d;      // same mapping as odeu_plan_create: one compartment = the single-compartment model
  if (dm.ode_id == ODEU_ODE_MULTI_HH && dm.num_compartments == 1) { dm.ode_id = ODEU_ODE_HODGKIN_HUXLEY; dm.num_compartments = 0; }
  plan.desc = dm;
  plan.theta_default.assign(theta_default, theta_default + p);
  switch (dm.ode_id) {
    case ODEU_ODE_LORENZ: return emu_solver<OdeLorenz>(plan, e, pf, g, s);
    case ODEU_ODE_VAN_DER_POL: return emu_solver<OdeVanDerPol>(plan, e, pf, g, s);
    case ODEU_ODE_LOTKA_VOLTERRA: return emu_solver<OdeLotkaVolterra>(plan, e, pf, g, s);
    case ODEU_ODE_PENDULUM: return emu_solver<OdePendulum>(plan, e, pf, g, s);
    case ODEU_ODE_LCAO: if (dm.ode_variant == 2) return emu_solver<OdeLCAO<2>>(plan, e, pf, g, s); break;
    case ODEU_ODE_HODGKIN_HUXLEY:
      if (dm.ode_variant == 0) return emu_solver<OdeHodgkinHuxley<0>>(plan, e, pf, g, s);
      if (dm.ode_variant == 1) return emu_solver<OdeHodgkinHuxley<1>>(plan, e, pf, g, s);
      if (dm.ode_variant == 4) return emu_solver<OdeHodgkinHuxley<4>>(plan, e, pf, g, s);
      break;
    case ODEU_ODE_MULTI_HH:
      if (dm.num_compartments == 2 && dm.ode_variant == 1) return emu_solver<OdeMultiHH<1, 2>>(plan, e, pf, g, s);
      if (dm.num_compartments == 2 && dm.ode_variant == 4) return emu_solver<OdeMultiHH<4, 2>>(plan, e, pf, g, s);
      break;
  }
  return -2;
}

// The gradient launcher normally launches a kernel; for the emulation the kernel launch is
// replaced by a host loop over (trajectory, chunk).  launch_grad's argument folding is reused by
// intercepting the arguments it builds: we re-run its validation through a host-only copy.
template <class Ode, class Tab>
int emu_grad(const odeu_plan& plan, const odeu_ekf_io& io, const odeu_grad_io& g) {
  GradArgs<Ode::NX, Ode::NP> a;
  if (int rc = fill_grad_args<Ode>(plan, io, &g, a)) return rc;
  if (rows_eligible<Ode, Tab, GDual<double, 1>>(io)) return emu_rows<Ode, Tab, GDual<double, 1>>(a, nullptr);
  using Cfg = GradCfg<Ode>;
  const int nchunks = (g.p_opt + Cfg::PC - 1) / Cfg::PC;
  for (int c = 0; c < nchunks; ++c)
    for (long long b = 0; b < io.B; ++b) ekf_grad_trajectory<Ode, Tab, Cfg::KC, Cfg::PC>(a, b, c);
  return 0;
}
}  // namespace odeu

extern "C" {
int hostemu_sens_run(const odeu_plan_desc* d, const double* theta_default, int p, const odeu_sens_io* s) {
  return odeu::emu_dispatch(*d, theta_default, p, nullptr, nullptr, nullptr, s);
}
int hostemu_grad_run(const odeu_plan_desc* d, const double* theta_default, int p, const odeu_ekf_io* io,
                     const odeu_grad_io* g) {
  return odeu::emu_dispatch(*d, theta_default, p, io, nullptr, g);
}
int hostemu_ekf_run(const odeu_plan_desc* d, const double* theta_default, int p, const odeu_ekf_io* io) {
  return odeu::emu_dispatch(*d, theta_default, p, io, nullptr);
}
int hostemu_pf_run(const odeu_plan_desc* d, const double* theta_default, int p, const odeu_pf_io* io) {
  return odeu::emu_dispatch(*d, theta_default, p, nullptr, io);
}
const char* hostemu_last_error() { return odeu::g_err.c_str(); }
}
