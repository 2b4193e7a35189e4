// Kernel instantiations: multi-compartment Hodgkin-Huxley chain
// (reference src/ode/hodgkin_huxley.py:284-439), 2 compartments (the shipped configs' shape).
#include "launch.cuh"
namespace odeu {
Launchers resolve_multi_hh(int model, int nc, int solver) {
  if (nc != 2) return {nullptr, nullptr, nullptr};
  switch (model) {
    case 1: return resolve_solver<OdeMultiHH<1, 2>>(solver);
    case 4: return resolve_solver<OdeMultiHH<4, 2>>(solver);
    default: return {nullptr, nullptr, nullptr};
  }
}
}
