// Kernel instantiations: Lorenz-63 (reference src/ode/lorenz.py) x all embedded RK tableaux.
#include "launch.cuh"
namespace odeu { Launchers resolve_lorenz(int solver) { return resolve_solver<OdeLorenz>(solver); } }
