// Host-side plan object behind the opaque `odeu_plan` of include/odeu.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/odeu.h"

struct odeu_plan {
  odeu_plan_desc desc;
  int n;   // flattened state dimension
  int p;   // number of scalar ODE parameters
  std::vector<double> theta_default;
  // launcher selected at plan creation (ode x solver)
  int (*ekf_launch)(const odeu_plan&, const odeu_ekf_io&, cudaStream_t);
  int (*pf_launch)(const odeu_plan&, const odeu_pf_io&, cudaStream_t);
  int (*rhs_launch)(const odeu_plan&, long long, double, const double*, const double*, const double*,
                    double*, cudaStream_t);
  // gradient run, or (sens != null) the parameter-sensitivity weights of the same ODE x solver
  int (*grad_launch)(const odeu_plan&, const odeu_ekf_io*, const odeu_grad_io*, const odeu_sens_io*, cudaStream_t);
  int (*rows_launch)(const odeu_plan&, const odeu_ekf_io&, cudaStream_t);   // null for small systems
};

namespace odeu {

void set_error(const char* fmt, ...);
void count_launch();

using EkfLaunchFn = int (*)(const odeu_plan&, const odeu_ekf_io&, cudaStream_t);
using PfLaunchFn = int (*)(const odeu_plan&, const odeu_pf_io&, cudaStream_t);
using RhsLaunchFn = int (*)(const odeu_plan&, long long, double, const double*, const double*,
                           const double*, double*, cudaStream_t);
using GradLaunchFn = int (*)(const odeu_plan&, const odeu_ekf_io*, const odeu_grad_io*, const odeu_sens_io*, cudaStream_t);
using RowsLaunchFn = int (*)(const odeu_plan&, const odeu_ekf_io&, cudaStream_t);
RowsLaunchFn resolve_rows_hh(int model, int solver);
RowsLaunchFn resolve_rows_multi_hh(int model, int nc, int solver);
GradLaunchFn resolve_grad_small(int ode_id, int variant, int solver);
GradLaunchFn resolve_grad_hh(int model, int solver);
GradLaunchFn resolve_grad_multi_hh(int model, int nc, int solver);
struct Launchers { EkfLaunchFn ekf; PfLaunchFn pf; RhsLaunchFn rhs; };

// One translation unit per ODE family instantiates its kernels and exposes a resolver.
Launchers resolve_lorenz(int solver);
Launchers resolve_van_der_pol(int solver);
Launchers resolve_lotka_volterra(int solver);
Launchers resolve_pendulum(int solver);
Launchers resolve_lcao(int D, int solver);
Launchers resolve_hh(int model, int solver);
Launchers resolve_multi_hh(int model, int nc, int solver);

}  // namespace odeu
