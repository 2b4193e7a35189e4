import functools

Partial = functools.partial
