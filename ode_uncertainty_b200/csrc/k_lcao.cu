// Kernel instantiations: linearly coupled anharmonic oscillators (reference src/ode/lcao.py),
// D = 2 (the reference's shape, n = 4).
#include "launch.cuh"
namespace odeu {
Launchers resolve_lcao(int D, int solver) {
  if (D == 2) return resolve_solver<OdeLCAO<2>>(solver);
  return {nullptr, nullptr, nullptr};
}
}
