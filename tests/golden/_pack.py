"""Compact container for many small named arrays: one blob per dtype + a JSON index, stored in a
single compressed .npz (a plain .npz costs a few hundred bytes of zip header per array)."""
import json

import numpy as np


class Packer:
    def __init__(self):
        self.blobs = {"int32": [], "int64": [], "float64": []}
        self.sizes = {k: 0 for k in self.blobs}
        self.index = {}

    def __setitem__(self, name, arr):
        a = np.asarray(arr)
        if a.dtype.kind == "f":
            a = a.astype(np.float64)
        elif a.dtype == np.int32:
            pass
        else:
            a = a.astype(np.int64)
        dt = str(a.dtype)
        flat = a.reshape(-1)
        self.index[name] = [dt, self.sizes[dt], int(flat.size), list(a.shape)]
        self.blobs[dt].append(flat)
        self.sizes[dt] += int(flat.size)

    def save(self, path):
        out = {"index": np.frombuffer(json.dumps(self.index).encode(), dtype=np.uint8)}
        for dt, parts in self.blobs.items():
            out[dt] = np.concatenate(parts) if parts else np.empty(0, dtype=dt)
        np.savez_compressed(path, **out)


class Pack:
    def __init__(self, path):
        z = np.load(path)
        self.index = json.loads(bytes(z["index"]).decode())
        self.blobs = {dt: z[dt] for dt in ("int32", "int64", "float64")}

    def __contains__(self, name):
        return name in self.index

    def __getitem__(self, name):
        dt, off, size, shape = self.index[name]
        return self.blobs[dt][off:off + size].reshape(shape).copy()
