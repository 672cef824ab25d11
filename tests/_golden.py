"""Loaders for tests/golden/*.npz (written by tests/golden/make_golden.py from the genuine reference)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from _pack import Pack  # noqa: E402
from oracle.oracle import Coo  # noqa: E402

_packs = {}


def pack(name) -> Pack:
    if name not in _packs:
        _packs[name] = Pack(os.path.join(HERE, "golden", name + ".npz"))
    return _packs[name]


def get_coo(p: Pack, prefix):
    if prefix + "_none" in p:
        return None
    shape = tuple(int(x) for x in p[prefix + "_shape"])
    idx = [p[f"{prefix}_idx{k}"] for k in range(len(shape))]
    so = tuple(int(x) for x in p[prefix + "_so"])
    return Coo(shape, idx, p[prefix + "_val"], None if so[0] < 0 else so)
