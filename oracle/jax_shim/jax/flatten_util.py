import torch


def ravel_pytree(tree):
    """Flatten a (nested) dict/list of arrays the way JAX does: dict keys in SORTED order."""
    leaves, rebuild = _flatten(tree)
    shapes = [l.shape for l in leaves]
    sizes = [l.numel() for l in leaves]
    flat = torch.cat([l.reshape(-1) for l in leaves]) if leaves else torch.zeros(0)

    def unravel(v):
        out, o = [], 0
        for shp, sz in zip(shapes, sizes):
            out.append(v[o:o + sz].reshape(shp))
            o += sz
        return rebuild(out)

    return flat, unravel


def _flatten(tree):
    if isinstance(tree, dict):
        keys = sorted(tree)
        subs = [_flatten(tree[k]) for k in keys]
        counts = [len(s[0]) for s in subs]
        leaves = [l for s in subs for l in s[0]]

        def rebuild(ls):
            out, o = {}, 0
            for k, s, c in zip(keys, subs, counts):
                out[k] = s[1](ls[o:o + c])
                o += c
            return out
        return leaves, rebuild
    if isinstance(tree, (list, tuple)):
        subs = [_flatten(t) for t in tree]
        counts = [len(s[0]) for s in subs]
        leaves = [l for s in subs for l in s[0]]

        def rebuild(ls):
            out, o = [], 0
            for s, c in zip(subs, counts):
                out.append(s[1](ls[o:o + c]))
                o += c
            return type(tree)(out)
        return leaves, rebuild
    t = torch.as_tensor(tree, dtype=torch.float64) if not isinstance(tree, torch.Tensor) else tree
    return [t], (lambda ls: ls[0])
