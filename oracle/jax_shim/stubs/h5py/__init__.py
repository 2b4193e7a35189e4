"""TEST INFRASTRUCTURE: a minimal stand-in for h5py (absent from the image) so that the reference's
own `scripts/run_filter.py::main` and `src/utils.py::store_data` run unmodified over the jax shim.
One "H5 file" is one .npz archive; only what those call sites use is provided
(File(path[, mode]) as a context manager, `f[name]`, `name in f.keys()`, `del f[name]`,
`f.create_dataset(name, data=...)`)."""
import os

import numpy as np


class File:
    def __init__(self, path, mode="r"):
        self.path, self.mode = str(path), mode
        self.data = {}
        if mode in ("r", "a", "r+") and os.path.exists(self.path):
            with np.load(self.path, allow_pickle=False) as f:
                self.data = {k: f[k] for k in f.files}
        elif mode == "r":
            raise FileNotFoundError(self.path)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        if self.mode != "r":
            with open(self.path, "wb") as fh:      # keep the caller's file name (np.savez appends .npz to str paths)
                np.savez(fh, **self.data)

    def keys(self):
        return self.data.keys()

    def __contains__(self, k):
        return k in self.data

    def __getitem__(self, k):
        return self.data[k]

    def __delitem__(self, k):
        del self.data[k]

    def create_dataset(self, name, data=None, **kw):
        if hasattr(data, "detach"):
            data = data.detach().cpu().numpy()
        self.data[name] = np.asarray(data)
        return self.data[name]
