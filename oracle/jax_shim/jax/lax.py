import builtins as _b

import torch
from torch.utils import _pytree as pt


def cond(pred, true_fn, false_fn, *operands):
    return true_fn(*operands) if bool(pred) else false_fn(*operands)


def scan(f, init, xs=None, length=None):
    if xs is None:
        n = length
    else:
        leaves = pt.tree_leaves(xs)
        n = leaves[0].shape[0]
    carry = init
    ys = []
    for i in range(n):
        x_i = None if xs is None else pt.tree_map(lambda a: a[i], xs)
        # JAX hands the body a FRESH pytree of tracers every call, so `state["y"] = ...` inside the body
        # (scripts/run_filter.py:206) never mutates the caller's initial_state nor an earlier output;
        # rebuild the containers to keep that purity
        carry, y = f(pt.tree_map(lambda a: a, carry), x_i)
        ys.append(y)
    if not ys or ys[0] is None:
        return carry, None
    stacked = pt.tree_map(lambda *a: torch.stack(a), *ys)
    return carry, stacked


def slice(x, start, limit, strides=None):
    strides = strides or (1,) * x.ndim
    idx = tuple(_b.slice(int(s), int(l), int(st)) for s, l, st in zip(start, limit, strides))
    return x[idx]


def while_loop(cond_fun, body_fun, init):
    v = init
    while bool(cond_fun(v)):
        v = body_fun(v)
    return v
