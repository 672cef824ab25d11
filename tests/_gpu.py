"""Helpers for the -m gpu parity tests: move oracle-side Coo objects through the C ABI."""
import numpy as np

import spsparse_b200 as sp
from oracle.oracle import Coo


def up(ctx, c: Coo):
    if c is None:
        return None
    return sp.CooArray.from_host(ctx, c.shape, c.idx, c.val, c.sort_order)


def down(a: sp.CooArray) -> Coo:
    idx, val = a.to_host()
    return Coo(a.shape, idx, val, a.sort_order)


class DevView:
    """Zero-copy view of library-owned device memory for torch.as_tensor."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def dev_to_numpy(ptr, n, typestr):
    import torch
    return torch.as_tensor(DevView(ptr, n, typestr), device="cuda").cpu().numpy()
