"""Host-side mirror of the reference's helper functions that sit on the hot path
(src/utils.py, SURVEY 8(a) A4/A6/A9 and the observation plumbing of A12).

Inside the fused kernels these operations never appear as separate calls (the covariance is carried
in full form); the functions below exist so code written against the
reference's `src.utils` keeps working at the API boundary.  They operate on torch tensors on
whatever device the inputs live on and are NOT part of the timed path.

Factor convention: the reference's QR-based factors are defined only up to column signs (SURVEY F1;
its own test compares `c @ c.T`, tests/test_utils.py:23-31).  Here the factor returned is the unique
lower-triangular one with a non-negative diagonal (the Cholesky factor of the sum), so
`c @ c.T` equals the reference's `c @ c.T`.
"""
from __future__ import annotations

import math
from typing import Dict, Union

import numpy as np
import torch

from .runners import isin_tolerance, sync_times  # noqa: F401  (src/utils.py:181-215)

Tensor = torch.Tensor


def _t(a) -> Tensor:
    return a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, dtype=np.float64))


def const_diag(n: int, val: float) -> Tensor:
    """src/utils.py:39-51."""
    return torch.diag(torch.full((n,), float(val), dtype=torch.float64))


def _factor_of_sum(*blocks: Tensor) -> Tensor:
    """Lower-triangular L with L L^T = sum_k b_k b_k^T, also for a singular sum (R-factor of the
    stacked transposed blocks like the reference, then signs fixed to a non-negative diagonal)."""
    stacked = torch.cat([_t(b).to(torch.float64).transpose(-1, -2) for b in blocks], dim=-2)
    R = torch.linalg.qr(stacked, mode="r")[1]
    sgn = torch.sign(torch.diagonal(R, dim1=-2, dim2=-1))
    sgn = torch.where(sgn == 0, torch.ones_like(sgn), sgn)
    return (R * sgn.unsqueeze(-1)).transpose(-1, -2)


def sqrt_L_sum_qr(a: Tensor, b: Tensor) -> Tensor:
    """src/utils.py:233-252: factor of a a^T + b b^T."""
    return _factor_of_sum(a, b)


def sqrt_L_sum_qr_3(a: Tensor, b: Tensor, c: Tensor) -> Tensor:
    """src/utils.py:255-274: factor of a a^T + b b^T + c c^T."""
    return _factor_of_sum(a, b, c)


def negative_log_gaussian_sqrt(x: Tensor, m: Tensor, P_sqrt: Tensor) -> Tensor:
    """src/utils.py:109-128 (uses |diag P_sqrt|, so it is insensitive to the factor's signs)."""
    x, m, P_sqrt = _t(x).to(torch.float64), _t(m).to(torch.float64), _t(P_sqrt).to(torch.float64)
    n = m.shape[-1]
    y = torch.linalg.solve_triangular(P_sqrt, (x - m).unsqueeze(-1), upper=False).squeeze(-1)
    return (0.5 * (y * y).sum(-1) + n / 2 * math.log(2 * math.pi)
            + torch.log(torch.abs(torch.diagonal(P_sqrt, dim1=-2, dim2=-1))).sum(-1))


def _flat(values: Union[Dict[str, np.ndarray], np.ndarray]):
    """jax.flatten_util.ravel_pytree: dict leaves in SORTED key order."""
    if isinstance(values, dict):
        keys = sorted(values)
        parts = [np.asarray(values[k], dtype=np.float64) for k in keys]
        flat = np.concatenate([p.reshape(-1) for p in parts]) if parts else np.zeros(0)

        def unravel(v):
            out, o = {}, 0
            for k, p in zip(keys, parts):
                out[k] = v[o:o + p.size].reshape(p.shape)
                o += p.size
            return out
        return flat, unravel
    arr = np.asarray(values, dtype=np.float64)
    return arr.reshape(-1), (lambda v: v.reshape(arr.shape))


def normalize(values, mins, maxs):
    """src/utils.py:131-154: (v - min) / (max - min), pytree in, pytree out."""
    v, unravel = _flat(values)
    lo, _ = _flat(mins)
    hi, _ = _flat(maxs)
    return unravel((v - lo) / (hi - lo))


def inv_normalize(values, mins, maxs):
    """src/utils.py:156-178: v (max - min) + min."""
    v, unravel = _flat(values)
    lo, _ = _flat(mins)
    hi, _ = _flat(maxs)
    return unravel(v * (hi - lo) + lo)
