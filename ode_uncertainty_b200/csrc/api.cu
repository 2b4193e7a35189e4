// extern "C" entry points of include/odeu.h.
#include <atomic>
#include <cstring>

#include "plan.h"
#include "odes.cuh"

namespace odeu {

static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// Reference default parameters in ODEBuilder.params order.
static bool default_params(const odeu_plan_desc& d, int* n, std::vector<double>* th) {
  switch (d.ode_id) {
    case ODEU_ODE_LORENZ: *n = 3; *th = {10.0, 8.0 / 3, 28.0}; return true;           // lorenz.py:12-16
    case ODEU_ODE_VAN_DER_POL: *n = 2; *th = {5.0}; return true;                       // van_der_pol.py:12
    case ODEU_ODE_LOTKA_VOLTERRA: *n = 2; *th = {1.5, 1.0, 3.0, 1.0}; return true;     // lotka_volterra.py:12-17
    case ODEU_ODE_PENDULUM: *n = 2; *th = {3.0}; return true;                          // pendulum.py:12
    case ODEU_ODE_LCAO: *n = 2 * d.ode_variant; *th = {1.0, 2.0, 0.5}; return true;    // lcao.py:12-16
    case ODEU_ODE_HODGKIN_HUXLEY: {                                                    // hodgkin_huxley.py:64-81
      *n = d.ode_variant == 0 ? 8 : d.ode_variant == 1 ? 7 : d.ode_variant == 4 ? 4 : -1;
      *th = {1.0, 8.3e-5, 25.0, 53.0, 7.0, -107.0, 0.1, -70.0, -60.0, 0.01, 4e3, 0.01, 120.0, 0.01, 2.0};
      return *n > 0;
    }
    case ODEU_ODE_MULTI_HH: {                                                          // hodgkin_huxley.py:287-306
      const int dim = d.ode_variant == 0 ? 8 : d.ode_variant == 1 ? 7 : d.ode_variant == 4 ? 4 : -1;
      if (dim < 0 || d.num_compartments != 2) return false;
      *n = 2 * dim;
      *th = {1.0,                 // coupling_coeffs
             1.0,                 // C
             4.15e-5, 4.15e-5,    // A
             25.0, 20.0,          // g_Na
             53.0, 53.0,          // E_Na
             7.0, 10.0,           // g_K
             -107.0, -107.0,      // E_K
             0.09, 0.11,          // g_leak
             -70.0, -70.0,        // E_leak
             -60.0, -60.0,        // V_T
             0.01, 0.01,          // g_M
             4e3, 4e3,            // tau_max
             0.01, 0.01,          // g_L
             120.0, 120.0,        // E_Ca
             0.01, 0.01,          // g_T
             2.0, 2.0};           // V_x
      return true;
    }
    default: return false;
  }
}

}  // namespace odeu

using namespace odeu;

// defined in k_lorenz.cu (needs the launcher header)
long long odeu_sched_bytes(int n, long long B, long long T);

extern "C" {

int odeu_version(void) { return ODEU_VERSION; }

int odeu_plan_create(const odeu_plan_desc* desc_in, odeu_plan** out) {
  if (!desc_in || !out) { set_error("odeu_plan_create: null argument"); return -1; }
  *out = nullptr;
  // MultiCompartmentHodgkinHuxley(num_compartments = 1) (hodgkin_huxley.py:284-439 with an empty coupling_coeffs) IS
  // the single-compartment model: same equations, and its flat parameter layout [C, A, g_Na, ..., V_x] is the
  // single-compartment order, so the plan is served by the single-compartment kernels
  odeu_plan_desc mapped = *desc_in;
  if (mapped.ode_id == ODEU_ODE_MULTI_HH && mapped.num_compartments == 1) {
    mapped.ode_id = ODEU_ODE_HODGKIN_HUXLEY;
    mapped.num_compartments = 0;
  }
  const odeu_plan_desc* desc = &mapped;
  if (!(desc->step_size > 0.0)) { set_error("odeu_plan_create: step_size must be > 0"); return -1; }
  if (desc->cov_fn_id < ODEU_COV_DIAGONAL || desc->cov_fn_id > ODEU_COV_STATIC_DIAGONAL) {
    set_error("odeu_plan_create: unknown cov_fn_id %d", desc->cov_fn_id);
    return -1;
  }
  odeu_plan* p = new odeu_plan();
  p->desc = *desc;
  if (!default_params(*desc, &p->n, &p->theta_default)) {
    set_error("odeu_plan_create: unsupported ODE (id=%d variant=%d compartments=%d)", desc->ode_id,
              desc->ode_variant, desc->num_compartments);
    delete p;
    return -2;
  }
  p->p = (int)p->theta_default.size();
  Launchers fn = {nullptr, nullptr, nullptr};
  switch (desc->ode_id) {
    case ODEU_ODE_LORENZ: fn = resolve_lorenz(desc->solver_id); break;
    case ODEU_ODE_VAN_DER_POL: fn = resolve_van_der_pol(desc->solver_id); break;
    case ODEU_ODE_LOTKA_VOLTERRA: fn = resolve_lotka_volterra(desc->solver_id); break;
    case ODEU_ODE_PENDULUM: fn = resolve_pendulum(desc->solver_id); break;
    case ODEU_ODE_LCAO: fn = resolve_lcao(desc->ode_variant, desc->solver_id); break;
    case ODEU_ODE_HODGKIN_HUXLEY: fn = resolve_hh(desc->ode_variant, desc->solver_id); break;
    case ODEU_ODE_MULTI_HH: fn = resolve_multi_hh(desc->ode_variant, desc->num_compartments, desc->solver_id); break;
    default: break;
  }
  const bool dense = desc->ode_id == ODEU_ODE_LCAO && desc->ode_variant >= 64 &&
                     (2 * desc->ode_variant) % 128 == 0 && 2 * desc->ode_variant <= 512 &&
                     desc->solver_id >= ODEU_SOLVER_RKF45 && desc->solver_id <= ODEU_SOLVER_HEUN_EULER;
  if (!fn.ekf && !dense) {
    set_error("odeu_plan_create: no kernel for ode=%d variant=%d compartments=%d solver=%d",
              desc->ode_id, desc->ode_variant, desc->num_compartments, desc->solver_id);
    delete p;
    return -2;
  }
  p->rows_launch = desc->ode_id == ODEU_ODE_HODGKIN_HUXLEY ? resolve_rows_hh(desc->ode_variant, desc->solver_id)
                   : desc->ode_id == ODEU_ODE_MULTI_HH
                       ? resolve_rows_multi_hh(desc->ode_variant, desc->num_compartments, desc->solver_id)
                       : nullptr;
  p->ekf_launch = fn.ekf;
  p->pf_launch = fn.pf;
  p->rhs_launch = fn.rhs;
  p->grad_launch = desc->ode_id == ODEU_ODE_HODGKIN_HUXLEY ? resolve_grad_hh(desc->ode_variant, desc->solver_id)
                   : desc->ode_id == ODEU_ODE_MULTI_HH
                       ? resolve_grad_multi_hh(desc->ode_variant, desc->num_compartments, desc->solver_id)
                       : resolve_grad_small(desc->ode_id, desc->ode_variant, desc->solver_id);
  *out = p;
  return 0;
}

void odeu_plan_destroy(odeu_plan* plan) { delete plan; }

int odeu_plan_state_dim(const odeu_plan* plan) { return plan ? plan->n : -1; }
int odeu_plan_num_params(const odeu_plan* plan) { return plan ? plan->p : -1; }
int odeu_plan_default_params(const odeu_plan* plan, double* out) {
  if (!plan || !out) { set_error("odeu_plan_default_params: null argument"); return -1; }
  std::memcpy(out, plan->theta_default.data(), sizeof(double) * plan->p);
  return 0;
}

int64_t odeu_ekf_workspace_bytes(const odeu_plan* plan, int64_t B, int64_t T) {
  if (!plan || B <= 0 || T <= 0) return 0;
  return odeu_sched_bytes(plan->n, (long long)B, (long long)T);
}

int odeu_ekf_run(const odeu_plan* plan, const odeu_ekf_io* io, void* cuda_stream) {
  if (!plan || !io) { set_error("odeu_ekf_run: null argument"); return -1; }
  if (!plan->ekf_launch) { set_error("odeu_ekf_run: this plan is served by odeu_ekf_dense_run"); return -2; }
  if (plan->rows_launch) {   // medium-size systems: row kernel when eligible
    const int rc = plan->rows_launch(*plan, *io, (cudaStream_t)cuda_stream);
    if (rc != -100) return rc;
  }
  return plan->ekf_launch(*plan, *io, (cudaStream_t)cuda_stream);
}

int odeu_pf_run(const odeu_plan* plan, const odeu_pf_io* io, void* cuda_stream) {
  if (!plan || !io) { set_error("odeu_pf_run: null argument"); return -1; }
  if (!plan->pf_launch) { set_error("odeu_pf_run: no ensemble kernel for this plan"); return -2; }
  return plan->pf_launch(*plan, *io, (cudaStream_t)cuda_stream);
}

int odeu_ekf_grad_run(const odeu_plan* plan, const odeu_ekf_io* io, const odeu_grad_io* grad,
                      void* cuda_stream) {
  if (!plan || !io || !grad) { set_error("odeu_ekf_grad_run: null argument"); return -1; }
  if (!plan->grad_launch) { set_error("odeu_ekf_grad_run: no gradient kernel for this plan"); return -2; }
  return plan->grad_launch(*plan, io, grad, nullptr, (cudaStream_t)cuda_stream);
}

int odeu_param_sensitivity(const odeu_plan* plan, const odeu_sens_io* io, void* cuda_stream) {
  if (!plan || !io) { set_error("odeu_param_sensitivity: null argument"); return -1; }
  if (!plan->grad_launch) { set_error("odeu_param_sensitivity: no gradient kernel for this plan"); return -2; }
  return plan->grad_launch(*plan, nullptr, nullptr, io, (cudaStream_t)cuda_stream);
}

int odeu_ode_rhs(const odeu_plan* plan, int64_t B, double t, const double* x, const double* theta,
                 const double* theta_shared, double* dx, void* cuda_stream) {
  if (!plan) { set_error("odeu_ode_rhs: null plan"); return -1; }
  if (!plan->rhs_launch) { set_error("odeu_ode_rhs: not available for this plan"); return -2; }
  return plan->rhs_launch(*plan, (long long)B, t, x, theta, theta_shared, dx, (cudaStream_t)cuda_stream);
}

int64_t odeu_launch_count(void) { return (int64_t)g_launches.load(); }

size_t odeu_last_error(char* buf, size_t buflen) {
  const std::string& e = g_last_error;
  if (buf && buflen) {
    const size_t k = e.size() < buflen - 1 ? e.size() : buflen - 1;
    std::memcpy(buf, e.data(), k);
    buf[k] = 0;
  }
  return e.size();
}

}  // extern "C"
