// Kernel instantiations: single-compartment Hodgkin-Huxley (reference src/ode/hodgkin_huxley.py:61-281),
// models full / reduced-1 / reduced-4.
#include "launch.cuh"
namespace odeu {
Launchers resolve_hh(int model, int solver) {
  switch (model) {
    case 0: return resolve_solver<OdeHodgkinHuxley<0>>(solver);
    case 1: return resolve_solver<OdeHodgkinHuxley<1>>(solver);
    case 4: return resolve_solver<OdeHodgkinHuxley<4>>(solver);
    default: return {nullptr, nullptr, nullptr};
  }
}
}
