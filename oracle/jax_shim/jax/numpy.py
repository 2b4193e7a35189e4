import builtins as _b
import math

import torch

pi = math.pi
float64 = torch.float64


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    return torch.as_tensor(x, dtype=dtype if dtype is not None else (torch.float64 if _is_floaty(x) else None))


def _is_floaty(x):
    if isinstance(x, float):
        return True
    if isinstance(x, (list, tuple)):
        return _b.any(_is_floaty(v) for v in x)
    return False


def _dt(dtype):
    if dtype is None:
        return None
    if dtype is float:
        return torch.float64
    if dtype is int:
        return torch.int64
    if dtype is bool:
        return torch.bool
    return dtype


def array(x, dtype=None):
    dt = _dt(dtype)
    if isinstance(x, torch.Tensor):
        return x.clone() if dt is None else x.to(dt)
    if isinstance(x, (list, tuple)) and len(x) and _b.any(isinstance(v, torch.Tensor) for v in x):
        return torch.stack([_t(v) for v in x])
    t = torch.as_tensor(x)
    if dt is not None:
        return t.to(dt)
    if t.dtype in (torch.float32, torch.float16):
        t = t.to(torch.float64)
    return t


asarray = array


def _shape(s):
    if isinstance(s, int):
        return (s,)
    if callable(s):      # `x.size` is an int property in JAX and a bound method on torch tensors
        n = 1            # (scripts/run_filter.py:75 `const_diag(x0_arr_built.size, 1e-12)`)
        for v in s():
            n *= int(v)
        return (n,)
    return tuple(int(v) for v in s)


def zeros(shape=(), dtype=None):
    return torch.zeros(_shape(shape), dtype=_dt(dtype) or torch.float64)


def ones(shape=(), dtype=None):
    return torch.ones(_shape(shape), dtype=_dt(dtype) or torch.float64)


def full(shape, val, dtype=None):
    if isinstance(val, bool) or (isinstance(val, torch.Tensor) and val.dtype == torch.bool):
        return torch.full(_shape(shape), bool(val), dtype=torch.bool)
    return torch.full(_shape(shape), float(val), dtype=_dt(dtype) or torch.float64)


def zeros_like(x):
    return torch.zeros_like(x)


def ones_like(x):
    return torch.ones_like(x)


def eye(n):
    return torch.eye(n)


def arange(*a, dtype=None):
    dt = _dt(dtype)
    if dt is None and _b.any(isinstance(v, float) for v in a):
        dt = torch.float64
    return torch.arange(*a, dtype=dt)


def diag(x, k=0):
    return torch.diag(x, k)


def outer(a, b):
    return torch.outer(a, b)


def stack(xs, axis=0):
    return torch.stack([_t(x) for x in xs], dim=axis)


def concatenate(xs, axis=0):
    return torch.cat(list(xs), dim=axis)


concat = concatenate


def split(x, n):
    return list(torch.tensor_split(x, n))


def broadcast_to(x, shape):
    return torch.broadcast_to(_t(x), _shape(shape))


def broadcast_shapes(*s):
    return tuple(torch.broadcast_shapes(*s))


def flip(x, axis=None):
    return torch.flip(x, dims=tuple(range(x.ndim)) if axis is None else (axis,))


def fill_diagonal(a, val, inplace=True):
    assert not inplace
    return a - torch.diag(torch.diagonal(a)) + torch.diag(torch.broadcast_to(_t(val), (a.shape[0],)))


exp = lambda x: torch.exp(_t(x))
log = lambda x: torch.log(_t(x))
log10 = lambda x: torch.log10(_t(x, torch.float64))
sin = lambda x: torch.sin(_t(x))
cos = lambda x: torch.cos(_t(x))
abs = lambda x: torch.abs(_t(x))
sqrt = lambda x: torch.sqrt(_t(x))
minimum = lambda a, b: torch.minimum(_t(a), _t(b))
logical_and = lambda a, b: torch.logical_and(_t(a), _t(b))
any = lambda x: torch.any(x)
all = lambda x: torch.all(x)
nan_to_num = lambda x: torch.nan_to_num(x)


def pow(a, b):
    return torch.pow(_t(a, torch.float64), _t(b, torch.float64))


def where(c, a, b):
    return torch.where(c, a, b)      # python scalars promote like jnp.where (ints stay ints)


def sum(x, axis=None):
    if axis is None:
        return torch.sum(x)
    if isinstance(axis, (list, tuple)) and len(axis) == 0:
        return x
    return torch.sum(x, dim=tuple(axis) if isinstance(axis, (list, tuple)) else axis)


def einsum(spec, *ops):
    return torch.einsum(spec, *ops)


def nonzero(x):
    return torch.nonzero(x, as_tuple=True)


def flatnonzero(x):
    return torch.nonzero(x.reshape(-1), as_tuple=True)[0]


def searchsorted(a, v):
    return torch.searchsorted(a, v)


def isin(a, b):
    return torch.isin(a, b)


class linalg:
    @staticmethod
    def norm(x):
        return torch.linalg.norm(x)

    @staticmethod
    def cholesky(x):
        return torch.linalg.cholesky(x)
