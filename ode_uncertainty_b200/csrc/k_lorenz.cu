// Kernel instantiations: Lorenz-63 (reference src/ode/lorenz.py) x all embedded RK tableaux.
#include "launch.cuh"
namespace odeu { Launchers resolve_lorenz(int solver) { return resolve_solver<OdeLorenz>(solver); } }

long long odeu_sched_bytes(int n, long long B, long long T) {
  odeu::SchedGeom g;
  return odeu::sched_geometry(n, B, T, g) ? g.bytes : 0;
}
