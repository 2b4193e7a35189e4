// Gradient-kernel instantiations: Lorenz, Van der Pol, Lotka-Volterra, pendulum, LCAO.
#include "launch_grad.cuh"
namespace odeu {
GradLaunchFn resolve_grad_small(int ode_id, int variant, int solver) {
  switch (ode_id) {
    case ODEU_ODE_LORENZ: return resolve_grad_solver<OdeLorenz>(solver);
    case ODEU_ODE_VAN_DER_POL: return resolve_grad_solver<OdeVanDerPol>(solver);
    case ODEU_ODE_LOTKA_VOLTERRA: return resolve_grad_solver<OdeLotkaVolterra>(solver);
    case ODEU_ODE_PENDULUM: return resolve_grad_solver<OdePendulum>(solver);
    case ODEU_ODE_LCAO: return variant == 2 ? resolve_grad_solver<OdeLCAO<2>>(solver) : nullptr;
    default: return nullptr;
  }
}
}
