import functools

from torch.utils import _pytree as pt


def map(f, tree, *rest, is_leaf=None):
    return pt.tree_map(f, tree, *rest, is_leaf=is_leaf)


def reduce(f, tree, *init):
    return functools.reduce(f, pt.tree_leaves(tree), *init)


def leaves(tree):
    return pt.tree_leaves(tree)


def structure(tree):
    return pt.tree_structure(tree)


def transpose(outer_treedef, inner_treedef, pytree_to_transpose):
    """dict-of-lists -> list-of-dicts (the only use: src/ode/hodgkin_huxley.py:418-431)."""
    assert isinstance(pytree_to_transpose, dict)
    keys = list(pytree_to_transpose)
    n = len(pytree_to_transpose[keys[0]])
    return [{k: pytree_to_transpose[k][i] for k in keys} for i in range(n)]
