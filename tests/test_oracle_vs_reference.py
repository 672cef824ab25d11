"""Live cross-check: restatement (oracle/spsparse_oracle.c) against the genuine reference
(oracle/_ref/libspsparse_ref.so) on fresh seeded inputs.  Skipped when oracle/_ref is absent."""
import numpy as np

import _cases
from oracle import oracle as O


def test_consolidate_live(orc, ref):
    for s in range(120, 260):
        c = _cases.consolidate_case(s)
        a = O.Coo(tuple(c["shape"]), c["idx"], c["val"])
        args = (a, tuple(c["sort_order"]), c["policy"], c["zero_nan"])
        g, w = orc.consolidate(*args), ref.consolidate(*args)
        assert _cases.same_coo(g, w), s
        assert np.array_equal(orc.sorted_permutation(a, args[1]), ref.sorted_permutation(a, args[1]))
        if a.rank == 2:
            assert np.array_equal(orc.dim_beginnings(g), ref.dim_beginnings(w))


def test_multiply_mm_live(orc, ref):
    for s in range(300, 700):
        c = _cases.mm_case(s, big=(s % 40 == 0))
        A, B = c["A"], c["B"]
        acon = orc.consolidate(A, (1, 0) if c["tA"] == "T" else (0, 1), c["policy"], c["zero_nan"])
        bcon = orc.consolidate(B, (0, 1) if c["tB"] == "T" else (1, 0), c["policy"], c["zero_nan"])
        if (A.n and not acon.n) or (B.n and not bcon.n):
            continue  # the reference segfaults here (see tests/golden/make_golden.py)
        args = (c["C"], c["si"], A, c["tA"], c["sj"], B, c["tB"], c["sk"], c["policy"], c["zero_nan"])
        assert _cases.same_coo(orc.multiply_mm(*args), ref.multiply_mm(*args)), s


def test_multiply_mv_live(orc, ref):
    for s in range(150, 400):
        c = _cases.mv_case(s)
        acon = orc.consolidate(c["A"], (1, 0) if c["tA"] == "T" else (0, 1), c["policy"], c["zero_nan"])
        if c["A"].n and not acon.n:
            continue  # reference UB: empty consolidated A (multiply_sparse.hpp:302-303 comment)
        args = (c["C"], c["si"], c["A"], c["tA"], c["sj"], c["V"], c["policy"], c["zero_nan"])
        assert _cases.same_coo(orc.multiply_mv(*args), ref.multiply_mv(*args)), s


def test_join_live(orc, ref):
    rng = np.random.default_rng(7)
    for _ in range(300):
        lists = [np.unique(rng.integers(0, 30, int(rng.integers(0, 20)))).astype(np.int32) for _ in range(3)]
        assert np.array_equal(orc.join(lists[0], lists[1]), ref.join(lists[0], lists[1]))
        assert np.array_equal(orc.join(*lists), ref.join(*lists))


def test_dense_ops_live(orc, ref):
    for s in range(100, 220):
        c = _cases.dense_case(s)
        a = O.Coo(tuple(c["shape"]), c["idx"], c["val"])
        assert _cases.same_coo(orc.transpose(a, tuple(c["perm"])), ref.transpose(a, tuple(c["perm"]))), s
        for policy in _cases.POLICIES:
            g, w = orc.to_dense(a, policy), ref.to_dense(a, policy)
            assert np.array_equal(g.view(np.uint64), w.view(np.uint64)), (s, policy)
        assert _cases.same_coo(orc.to_sparse(g), ref.to_sparse(w)), s
