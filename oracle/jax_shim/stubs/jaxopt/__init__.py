"""TEST INFRASTRUCTURE: stand-in for jaxopt.ScipyBoundedMinimize (jaxopt 0.8.3 is absent from the
image) so that the reference's own `optimize()` / `optimize_run()` (scripts/run_parameter_estimation.py:
49-308, 540-682) run unmodified over the jax shim.  Like jaxopt it flattens the parameter pytree,
hands SciPy's L-BFGS-B the value and gradient of `fun` (reverse mode through `jax.value_and_grad`
unless `value_and_grad=True`), and reports fun_val / iter_num / num_fun_eval / num_jac_eval."""
from collections import namedtuple

import numpy as np
import scipy.optimize
import torch

import jax
from jax.flatten_util import ravel_pytree

ScipyMinimizeInfo = namedtuple("ScipyMinimizeInfo", "fun_val success status iter_num hess_inv num_fun_eval num_jac_eval num_hess_eval")


class ScipyBoundedMinimize:
    def __init__(self, fun, method="L-BFGS-B", maxiter=500, jit=True, value_and_grad=False, tol=None, options=None, **_):
        self.fun, self.method, self.maxiter, self.value_and_grad, self.tol = fun, method, maxiter, value_and_grad, tol
        self.options = dict(options or {})

    def run(self, init_params, bounds, *args, **kwargs):
        z0, unravel = ravel_pytree(init_params)
        lo, _ = ravel_pytree(bounds[0])
        hi, _ = ravel_pytree(bounds[1])
        vg = self.fun if self.value_and_grad else jax.value_and_grad(self.fun)

        def f(z):
            v, g = vg(unravel(torch.as_tensor(np.asarray(z, dtype=np.float64))), *args, **kwargs)
            gf, _ = ravel_pytree(g)
            return float(v), np.asarray(gf.detach().numpy() if hasattr(gf, "detach") else gf, dtype=np.float64)

        res = scipy.optimize.minimize(f, np.asarray(z0.detach().numpy(), dtype=np.float64), jac=True, method=self.method,
                                      bounds=list(zip(lo.numpy(), hi.numpy())), tol=self.tol,
                                      options={**self.options, "maxiter": self.maxiter})
        info = ScipyMinimizeInfo(fun_val=torch.as_tensor(float(res.fun)), success=res.success, status=res.status,
                                 iter_num=torch.as_tensor(int(res.nit)), hess_inv=None, num_fun_eval=torch.as_tensor(int(res.nfev)),
                                 num_jac_eval=torch.as_tensor(int(res.njev)), num_hess_eval=torch.as_tensor(0))
        return unravel(torch.as_tensor(res.x)), info
