from . import tree_utils  # noqa: F401
