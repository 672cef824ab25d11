"""-m gpu: copy / transpose / to_dense / to_sparse through the C ABI against the fixtures the
genuine reference produced (tests/golden/dense_ops_cases.npz) and against the oracle on fresh seeds.  Dense
arrays are compared bit for bit except that any NaN equals any NaN (the GPU canonicalises NaN payloads)."""
import numpy as np
import pytest

import _cases
import _golden
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import spsparse_b200 as sp
    with sp.Context(0) as c:
        yield c


@pytest.fixture(scope="module")
def orc():
    return O.port()


def same_dense(a, b):
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(na, nb) and np.array_equal(a[~na].view(np.uint64), b[~nb].view(np.uint64)))


def test_dense_ops_fixtures_from_the_reference(ctx):
    import spsparse_b200 as sp
    from _gpu import up, down
    p = _golden.pack("dense_ops_cases")
    for s in range(int(p["count"])):
        a = _golden.get_coo(p, f"d{s}_in")
        args = [int(x) for x in p[f"d{s}_args"]]
        perm, policy = tuple(args[:-1]), args[-1]
        dA = up(ctx, a)
        T, Cp = sp.transpose(ctx, dA, perm), sp.copy(ctx, dA)
        assert _cases.same_coo(down(T), _golden.get_coo(p, f"d{s}_T")), s
        assert _cases.same_coo(down(Cp), a), s
        dense = sp.to_dense(ctx, dA, policy)
        assert same_dense(dense, p[f"d{s}_dense"]), s
        S = sp.to_sparse(ctx, p[f"d{s}_dense"])
        assert _cases.same_coo(down(S), _golden.get_coo(p, f"d{s}_sparse")), s
        for h in (dA, T, Cp, S):
            h.free()


def test_dense_ops_against_oracle_fresh_seeds(ctx, orc):
    import spsparse_b200 as sp
    from _gpu import up, down
    for s in range(300, 360):
        c = _cases.dense_case(s)
        a = O.Coo(tuple(c["shape"]), c["idx"], c["val"])
        dA = up(ctx, a)
        for policy in _cases.POLICIES:
            assert same_dense(sp.to_dense(ctx, dA, policy), orc.to_dense(a, policy)), (s, policy)
        # dense -> sparse -> dense is the identity on the dense side (storage order, zeros dropped)
        dense = orc.to_dense(a, c["policy"])
        S = sp.to_sparse(ctx, dense)
        assert same_dense(sp.to_dense(ctx, S), np.where(np.isnan(dense), dense, dense + 0.0))
        T = sp.transpose(ctx, dA, tuple(c["perm"]))
        assert _cases.same_coo(down(T), orc.transpose(a, tuple(c["perm"]))), s
        for h in (dA, S, T):
            h.free()


def test_dense_ops_larger_and_errors(ctx, orc):
    import spsparse_b200 as sp
    from _gpu import up
    rng = np.random.default_rng(3)
    shape, n = (1500, 2100), 2_000_000   # ~47 % of the cells hit, long duplicate runs on a few hot cells
    i, k = rng.integers(0, shape[0], n), rng.integers(0, shape[1], n)
    i[:50000], k[:50000] = 7, 11
    a = O.Coo(shape, [i, k], rng.integers(-3, 4, n).astype(np.float64))
    dA = up(ctx, a)
    for policy in _cases.POLICIES:
        assert same_dense(sp.to_dense(ctx, dA, policy), orc.to_dense(a, policy)), policy
    dA.free()
    bad = sp.CooArray.from_host(ctx, (4, 4), [[0, 5], [0, 0]], [1.0, 2.0])
    with pytest.raises(sp.SpbError):
        sp.to_dense(ctx, bad)   # index outside the shape (VectorCooArray::add would have refused it, :245-262)
    with pytest.raises(sp.SpbError):
        sp.transpose(ctx, bad, (0, 0))
    bad.free()


def test_netcdf_round_trip_through_the_python_mirror(ctx, tmp_path):
    """CooArray.to_netcdf / from_netcdf (spsparse_b200/ncio.py: the layout of ncio_spsparse, netcdf.hpp:86-138): the array that
    comes back consolidates to the same result, entry for entry."""
    import spsparse_b200 as sp
    rng = np.random.default_rng(12)
    n = 5000
    idx = [rng.integers(0, 700, n), rng.integers(0, 100000, n)]
    val = rng.standard_normal(n)
    A = sp.CooArray.from_host(ctx, (700, 100000), idx, val)
    p = str(tmp_path / "a.nc")
    A.to_netcdf(p, "A")
    B = sp.CooArray.from_netcdf(ctx, p, "A")
    (bi, bv) = B.to_host()
    assert np.array_equal(bi[0], idx[0]) and np.array_equal(bi[1], idx[1]) and np.array_equal(bv, val)
    Ra, Rb = sp.consolidate(ctx, A, sp.ROW_MAJOR), sp.consolidate(ctx, B, sp.ROW_MAJOR)
    (ia, va), (ib, vb) = Ra.to_host(), Rb.to_host()
    assert all(np.array_equal(x, y) for x, y in zip(ia, ib)) and np.array_equal(va, vb)
    for x in (A, B, Ra, Rb):
        x.free()
