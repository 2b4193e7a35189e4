// Gradient / row kernel instantiations: 2-compartment Hodgkin-Huxley (BASELINE config 3).
#include "launch_grad.cuh"
namespace odeu {
GradLaunchFn resolve_grad_multi_hh(int model, int nc, int solver) {
  if (nc != 2) return nullptr;
  switch (model) {
    case 1: return resolve_grad_solver<OdeMultiHH<1, 2>>(solver);
    case 4: return resolve_grad_solver<OdeMultiHH<4, 2>>(solver);
    default: return nullptr;
  }
}
RowsLaunchFn resolve_rows_multi_hh(int model, int nc, int solver) {
  if (nc != 2) return nullptr;
  switch (model) {
    case 1: return resolve_rows_solver<OdeMultiHH<1, 2>>(solver);
    case 4: return resolve_rows_solver<OdeMultiHH<4, 2>>(solver);
    default: return nullptr;
  }
}
}  // namespace odeu
