"""ORACLE tooling (test infrastructure only): a minimal `jax` look-alike on torch (CPU, float64).

Purpose: execute the reference's OWN source files (`/root/reference/src/**`, and the loop owners
in `/root/reference/scripts/*.py`) unmodified in a container that has no JAX, so that golden
vectors for the EKF path come from the reference code itself rather than from a restatement
(oracle/make_golden_ref.py).  Only the API surface those files touch is provided:
jnp/jsp/lax/tree/vmap/jvp/jacfwd/grad/jit/flatten_util; transformations map onto torch.func.
This is not JAX: XLA's fusion/rounding order is not reproduced (differences are at the 1e-16
level), PRNG functions are absent.
"""
from __future__ import annotations

import functools

import torch
from torch.func import grad as _tgrad
from torch.func import jacfwd as _tjacfwd
from torch.func import jvp as _tjvp
from torch.func import vjp as _tvjp
from torch.func import vmap as _tvmap

torch.set_default_dtype(torch.float64)
Array = torch.Tensor


# ---- x.at[idx].set/add/get -------------------------------------------------------------------
class _At:
    def __init__(self, x):
        self.x = x

    def __getitem__(self, idx):
        return _AtIdx(self.x, idx)


def _norm_idx(idx):
    def one(i):
        if isinstance(i, torch.Tensor) and i.dtype in (torch.int64, torch.int32, torch.bool) and i.ndim == 0:
            return int(i)
        return i
    if isinstance(idx, tuple):
        return tuple(one(i) for i in idx)
    return one(idx)


class _AtIdx:
    def __init__(self, x, idx):
        self.x, self.idx = x, _norm_idx(idx)

    def _put(self, v, accumulate):
        x = self.x
        lin = torch.arange(x.numel()).reshape(x.shape)[self.idx]
        v = torch.as_tensor(v, dtype=x.dtype)
        vb = torch.broadcast_to(v, lin.shape).reshape(-1)
        flat = x.reshape(-1)
        out = torch.index_put(flat, (lin.reshape(-1),), vb, accumulate=accumulate)
        return out.reshape(x.shape)

    def set(self, v, **_):
        if self.x.dtype == torch.bool:
            out = self.x.clone()
            out[self.idx] = v
            return out
        return self._put(v, False)

    def add(self, v, **_):
        return self._put(v, True)

    def get(self, **_):
        return self.x[self.idx]


torch.Tensor.at = property(lambda self: _At(self))
# jnp spells the keyword arguments of .diagonal() axis1/axis2 (src/utils.py:127)
_orig_diagonal = torch.Tensor.diagonal


def _diagonal(self, offset=0, dim1=0, dim2=1, axis1=None, axis2=None):
    return _orig_diagonal(self, offset=offset, dim1=dim1 if axis1 is None else axis1,
                          dim2=dim2 if axis2 is None else axis2)


torch.Tensor.diagonal = _diagonal


# ---- transformations -------------------------------------------------------------------------
def jit(f=None, static_argnums=None, static_argnames=None, **_):
    if f is None:
        return lambda g: g
    return f


def vmap(f, in_axes=0, out_axes=0):
    return _tvmap(f, in_dims=in_axes, out_dims=out_axes)


def jvp(f, primals, tangents, has_aux=False):
    return _tjvp(f, tuple(primals), tuple(tangents), has_aux=has_aux)


def vjp(f, *primals, has_aux=False):
    return _tvjp(f, *primals, has_aux=has_aux)


def jacfwd(f, argnums=0, has_aux=False):
    return _tjacfwd(f, argnums=argnums, has_aux=has_aux)


def grad(f, argnums=0, has_aux=False):
    return _tgrad(f, argnums=argnums, has_aux=has_aux)


def value_and_grad(f, argnums=0, has_aux=False):
    def wrapped(*a, **k):
        def g(*aa):
            out = f(*aa, **k)
            return (out[0], out) if has_aux else (out, out)
        gr, val = _tgrad(g, argnums=argnums, has_aux=True)(*a)
        return val, gr
    return wrapped


def clear_caches():
    pass


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


class _Debug:
    @staticmethod
    def print(*a, **k):
        pass


debug = _Debug()

from . import flatten_util, lax, numpy, random, scipy, tree, tree_util  # noqa: E402,F401
