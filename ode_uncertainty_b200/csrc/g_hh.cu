// Gradient / row kernel instantiations: single-compartment Hodgkin-Huxley.
#include "launch_grad.cuh"
namespace odeu {
GradLaunchFn resolve_grad_hh(int model, int solver) {
  switch (model) {
    case 0: return resolve_grad_solver<OdeHodgkinHuxley<0>>(solver);
    case 1: return resolve_grad_solver<OdeHodgkinHuxley<1>>(solver);
    case 4: return resolve_grad_solver<OdeHodgkinHuxley<4>>(solver);
    default: return nullptr;
  }
}
RowsLaunchFn resolve_rows_hh(int model, int solver) {
  switch (model) {
    case 0: return resolve_rows_solver<OdeHodgkinHuxley<0>>(solver);
    case 1: return resolve_rows_solver<OdeHodgkinHuxley<1>>(solver);
    default: return nullptr;   // reduced-4 (n = 4) stays on the register kernel
  }
}
}  // namespace odeu
