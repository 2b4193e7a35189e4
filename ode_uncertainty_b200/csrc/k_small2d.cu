// Kernel instantiations: the 2-state plugins (Van der Pol, Lotka-Volterra, pendulum;
// reference src/ode/{van_der_pol,lotka_volterra,pendulum}.py) x all embedded RK tableaux.
#include "launch.cuh"
namespace odeu {
Launchers resolve_van_der_pol(int solver) { return resolve_solver<OdeVanDerPol>(solver); }
Launchers resolve_lotka_volterra(int solver) { return resolve_solver<OdeLotkaVolterra>(solver); }
Launchers resolve_pendulum(int solver) { return resolve_solver<OdePendulum>(solver); }
}
