"""Seeded random inputs shared by the oracle tests, the golden-fixture generator and the GPU
parity tests.  Every case is a plain dict of numpy arrays / scalars so it can be stored in .npz."""
from __future__ import annotations

import numpy as np

from oracle.oracle import ADD, LEAVE_ALONE, REPLACE, Coo

POLICIES = (LEAVE_ALONE, ADD, REPLACE)


def _values(rng, n, flavour):
    """flavour: 'pos' strictly positive; 'int' small signed integers (exact in any order, cancels);
    'mixed' signed reals with zeros (+0/-0) and NaNs sprinkled in."""
    if flavour == "pos":
        return 0.5 + rng.random(n)
    if flavour == "int":
        return rng.integers(-3, 4, n).astype(np.float64)
    v = rng.standard_normal(n)
    if n:
        k = rng.random(n)
        v[k < 0.10] = 0.0
        v[(k >= 0.10) & (k < 0.14)] = -0.0
        v[(k >= 0.14) & (k < 0.22)] = np.nan
    return v


def consolidate_case(seed):
    rng = np.random.default_rng(1000 + seed)
    rank = 1 if seed % 7 == 3 else 2
    n = int([0, 1, 2, 5, 33, 100, 257, 1000, 2500, 4099][seed % 10])
    if rank == 2:
        shape = [(3, 4), (40, 50), (1 << 12, 1 << 11), (5, 100000), (100000, 7)][seed % 5]
    else:
        shape = [(6,), (500,), (1 << 20,)][seed % 3]
    idx = [rng.integers(0, s, n).astype(np.int32) for s in shape]
    flavour = ["pos", "int", "mixed"][seed % 3]
    val = _values(rng, n, flavour)
    so = (0,) if rank == 1 else ((0, 1) if (seed // 2) % 2 == 0 else (1, 0))
    return dict(shape=np.array(shape, np.int64), idx=idx, val=val, sort_order=np.array(so, np.int32),
                policy=POLICIES[seed % 3], zero_nan=int((seed // 3) % 2))


def dense_case(seed):
    """Input for transpose / to_dense / to_sparse: small shapes (the dense form is materialised), many
    duplicate tuples, all value flavours."""
    rng = np.random.default_rng(5000 + seed)
    rank = 1 if seed % 5 == 4 else 2
    n = int([0, 1, 7, 60, 400, 3000][seed % 6])
    shape = [(3, 4), (40, 50), (64, 33), (1, 200), (257, 2)][seed % 5] if rank == 2 else [(6,), (500,)][seed % 2]
    idx = [rng.integers(0, s, n).astype(np.int32) for s in shape]
    val = _values(rng, n, ["pos", "int", "mixed"][seed % 3])
    perm = (0,) if rank == 1 else ((1, 0) if seed % 2 == 0 else (0, 1))
    return dict(shape=np.array(shape, np.int64), idx=idx, val=val, perm=np.array(perm, np.int32), policy=POLICIES[seed % 3])


def _sparse_vec(rng, dim, flavour, extra_shape=0):
    """Scale vector as the reference wants it: ascending, unique; explicit zeros allowed."""
    k = int(rng.integers(1, dim + 1))
    ix = np.sort(rng.choice(dim, size=k, replace=False)).astype(np.int32)
    v = _values(rng, k, "pos" if flavour == "pos" else "int")
    if flavour != "pos" and k:
        v[rng.random(k) < 0.15] = 0.0
    return Coo((dim + extra_shape,), [ix], v, (0,))


def mm_case(seed, big=False):
    rng = np.random.default_rng(5000 + seed)
    if big:
        m, nj, nk = (int(x) for x in rng.integers(50, 400, 3))
        dens = 0.02 + 0.1 * rng.random()
    else:
        m, nj, nk = (int(x) for x in rng.integers(1, 9, 3))
        dens = rng.random()
    tA = "T" if seed % 2 else "."
    tB = "T" if (seed // 2) % 2 else "."
    flavour = ["pos", "int", "mixed"][seed % 3]
    shpA = (nj, m) if tA == "T" else (m, nj)
    shpB = (nk, nj) if tB == "T" else (nj, nk)

    def mat(shape):
        n = int(dens * shape[0] * shape[1] * 1.5) + int(rng.integers(0, 3))
        return Coo(shape, [rng.integers(0, shape[0], n), rng.integers(0, shape[1], n)],
                   _values(rng, n, flavour), None)

    A, B = mat(shpA), mat(shpB)
    use = rng.random(3) < 0.5
    si = _sparse_vec(rng, m, flavour, int(rng.integers(0, 3))) if use[0] else None
    sj = _sparse_vec(rng, nj, flavour) if use[1] else None
    sk = _sparse_vec(rng, nk, flavour, int(rng.integers(0, 3))) if use[2] else None
    Cst = [1.0, 2.5, -3.0][seed % 3]
    return dict(C=Cst, si=si, A=A, tA=tA, sj=sj, B=B, tB=tB, sk=sk, policy=POLICIES[(seed // 3) % 3],
                zero_nan=int((seed // 5) % 2))


def mv_case(seed):
    rng = np.random.default_rng(9000 + seed)
    m, nj = (int(x) for x in rng.integers(1, 12, 2))
    tA = "T" if seed % 2 else "."
    flavour = ["pos", "int", "mixed"][seed % 3]
    shpA = (nj, m) if tA == "T" else (m, nj)
    n = int(rng.random() * shpA[0] * shpA[1] * 1.5)
    A = Coo(shpA, [rng.integers(0, shpA[0], n), rng.integers(0, shpA[1], n)], _values(rng, n, flavour), None)
    nv = int(rng.integers(0, nj + 3))
    V = Coo((nj,), [rng.integers(0, nj, nv)], _values(rng, nv, flavour), None)
    use = rng.random(2) < 0.5
    si = _sparse_vec(rng, m, flavour) if use[0] else None
    sj = _sparse_vec(rng, nj, flavour) if use[1] else None
    return dict(C=[1.0, -0.5][seed % 2], si=si, A=A, tA=tA, sj=sj, V=V, policy=POLICIES[(seed // 3) % 3],
                zero_nan=int((seed // 5) % 2))


def same_coo(a: Coo, b: Coo, rtol=0.0):
    """Structure bit-exact; values bit-exact (rtol=0, NaN==NaN) or within rtol relative."""
    if tuple(a.shape) != tuple(b.shape) or a.n != b.n:
        return False
    for x, y in zip(a.idx, b.idx):
        if not np.array_equal(x, y):
            return False
    if rtol == 0.0:
        return bool(np.array_equal(a.val.view(np.uint64), b.val.view(np.uint64)) or
                    np.array_equal(a.val, b.val, equal_nan=True) and
                    np.array_equal(np.signbit(a.val), np.signbit(b.val)))
    nan_a, nan_b = np.isnan(a.val), np.isnan(b.val)
    if not np.array_equal(nan_a, nan_b):
        return False
    ok = ~nan_a
    return bool(np.all(np.abs(a.val[ok] - b.val[ok]) <= rtol * np.abs(b.val[ok])))
