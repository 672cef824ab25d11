"""Pins the CPU oracle (oracle/spsparse_oracle.c) to the reference: (1) the known answers written
in the reference's own tests, (2) fixtures produced by the genuine reference (tests/golden)."""
import numpy as np
import pytest

import _cases
import _golden
from oracle import oracle as O


# ---- tests/test_array.cpp:67-79
def test_sorted_permutation_known_answer(orc):
    a = O.Coo((2, 4), [[1, 1, 0], [3, 2, 3]], [5., 3., 17.])
    assert orc.sorted_permutation(a, (0, 1)).tolist() == [2, 1, 0]
    assert orc.sorted_permutation(a, (1, 0)).tolist() == [1, 2, 0]


# ---- tests/test_array.cpp:135-168
def test_consolidate_known_answer(orc):
    a = O.Coo((2, 4), [[1, 1, 0, 0, 1], [3, 2, 3, 1, 2]], [5., 3., 17., 14., 15.])
    r = orc.consolidate(a, (0, 1))
    assert r.idx[0].tolist() == [0, 0, 1, 1] and r.idx[1].tolist() == [1, 3, 2, 3]
    assert r.val.tolist() == [14., 17., 18., 5.]
    assert orc.dim_beginnings(r).tolist() == [0, 2, 4]
    r = orc.consolidate(a, (1, 0))  # entries reordered, index columns NOT swapped
    assert r.idx[0].tolist() == [0, 1, 0, 1] and r.idx[1].tolist() == [1, 2, 3, 3]
    assert r.val.tolist() == [14., 18., 17., 5.]
    assert orc.dim_beginnings(r).tolist() == [0, 1, 2, 4]


# ---- tests/test_array.cpp:170-218 (row walk over the compressed row starts)
def test_dim_beginnings_rows(orc):
    a = O.Coo((20, 10), [[1, 1, 2, 6], [0, 3, 4, 4]], [15., 17., 17., 10.])
    r = orc.consolidate(a, (0, 1))
    db = orc.dim_beginnings(r)
    assert db.tolist() == [0, 2, 3, 4]
    assert [int(r.idx[0][s]) for s in db[:-1]] == [1, 2, 6]
    assert r.idx[1][db[0]:db[1]].tolist() == [0, 3] and r.val[db[0]:db[1]].tolist() == [15., 17.]


# ---- tests/test_xiter.cpp:52-125
def test_join_known_answers(orc):
    assert orc.join([0, 2, 4, 6], [0, 1, 2, 3, 4, 5, 6, 7]).tolist() == [0, 2, 4, 6]
    assert orc.join([0, 1, 2, 3, 4, 5, 6, 7], [0, 2, 4, 6]).tolist() == [0, 2, 4, 6]
    assert orc.join([0, 2, 4, 5, 6, 7, 8, 9], [1, 2, 3, 4, 6]).tolist() == [2, 4, 6]
    assert orc.join([0, 2, 4, 6], [0, 1, 2, 3, 4, 5, 6, 7], [1, 2, 3, 6]).tolist() == [2, 6]
    assert orc.join([], [1, 2]).tolist() == [] and orc.join([1, 2], []).tolist() == []


# ---- tests/test_multiply_sparse.cpp:41-79 (disabled in the reference; verified against it)
@pytest.mark.parametrize("pairs", [False, True])
def test_multiply_scaled_known_answer(orc, pairs):
    row = O.Coo((2, 10), [[0, 0, 0, 0, 1], [8, 4, 0, 3, 8]], [6., 4., 2., 3., 3.])
    scale = O.Coo((10,), [[0, 4, 8]], [2., 4., 4.], (0,))
    col = O.Coo((10, 1), [[0, 3, 8], [0, 0, 0]], [2., 3., 5.])
    eye = O.Coo((10,), [np.arange(10)], np.ones(10), (0,))
    r = orc.multiply_mm(1.0, eye, row, ".", scale, col, ".", eye, pairs=pairs)
    assert r.shape == (2, 1)
    assert r.idx[0].tolist() == [0, 1] and r.idx[1].tolist() == [0, 0] and r.val.tolist() == [128., 60.]


# ---- SURVEY 8(c) quirks, verified on the genuine reference
def test_quirks(orc):
    # (1) zeros dropped on input, cancelling duplicates keep an explicit 0.0
    r = orc.consolidate(O.Coo((3, 3), [[1, 1, 2], [1, 1, 2]], [1., -1., 0.]), (0, 1))
    assert r.n == 1 and r.val.tolist() == [0.0]
    # (2) zero_nan drops NaN only in the leading run of the sorted sequence
    a = O.Coo((4, 4), [[0, 1, 2], [0, 1, 2]], [np.nan, 1., np.nan])
    r = orc.consolidate(a, (0, 1), zero_nan=True)
    assert r.idx[0].tolist() == [1, 2] and np.isnan(r.val[1])
    # (4) multiply drops exact-zero dot products, keeps NaN
    A = O.Coo((1, 2), [[0, 0], [0, 1]], [1., 1.])
    B = O.Coo((2, 2), [[0, 1, 0], [0, 0, 1]], [1., -1., np.nan])
    r = orc.multiply_mm(1.0, None, A, ".", None, B, ".", None)
    assert r.idx[1].tolist() == [1] and np.isnan(r.val[0])
    # inner-dimension mismatch is an error (multiply_sparse.hpp:172-174)
    with pytest.raises(O.InnerDimError):
        orc.multiply_mm(1.0, None, O.Coo((2, 3), [[0], [0]], [1.]), ".", None, O.Coo((2, 2), [[0], [0]], [1.]), ".", None)


# ---- tests/test_multiply_sparse.cpp:84-136 and :138-203, seeds 1..999, inputs+outputs from the reference
def test_reference_random_tests(orc):
    p = _golden.pack("reference_random_tests")
    eye = O.Coo((5,), [np.arange(5)], np.ones(5), (0,))
    for seed in range(1, 1000):
        A, B, Cg = (_golden.get_coo(p, f"mm{seed}_{x}") for x in "ABC")
        for pairs in (False, True):
            got = orc.multiply_mm(1.0, None, A, ".", eye, B, ".", None, pairs=pairs)
            assert _cases.same_coo(got, Cg), f"MM seed {seed} pairs={pairs}"
        # the reference's own acceptance test: dense triple loop, 4 ulp (:112-126)
        Ad = np.zeros((5, 5)); np.add.at(Ad, (A.idx[0], A.idx[1]), A.val)
        Bd = np.zeros((5, 5)); np.add.at(Bd, (B.idx[0], B.idx[1]), B.val)
        Cd = np.zeros((5, 5)); Cd[got.idx[0], got.idx[1]] = got.val
        dense = np.zeros((5, 5))
        for k in range(5):
            dense += np.outer(Ad[:, k], Bd[k, :])
        assert np.all(np.abs(dense - Cd) <= 4 * np.spacing(np.maximum(np.abs(dense), np.abs(Cd))))
        A, V, Cg = (_golden.get_coo(p, f"mv{seed}_{x}") for x in "AVC")
        assert _cases.same_coo(orc.multiply_mv(1.0, None, A, ".", None, V), Cg), f"MV seed {seed}"


def test_consolidate_fixtures(orc):
    p = _golden.pack("consolidate_cases")
    for s in range(int(p["count"])):
        a, want = _golden.get_coo(p, f"c{s}_in"), _golden.get_coo(p, f"c{s}_out")
        pol, zn, *so = (int(x) for x in p[f"c{s}_args"])
        got = orc.consolidate(a, tuple(so), pol, zn)
        assert _cases.same_coo(got, want), f"consolidate case {s}"
        if a.rank == 2:
            assert np.array_equal(orc.dim_beginnings(got), p[f"c{s}_db"]), f"dim_beginnings case {s}"


def test_multiply_mm_fixtures(orc):
    p = _golden.pack("multiply_mm_cases")
    for s in range(int(p["count"])):
        si, A, sj, B, sk, want = (_golden.get_coo(p, f"m{s}_{x}") for x in ("si", "A", "sj", "B", "sk", "out"))
        Cst, tA, tB, pol, zn = p[f"m{s}_args"]
        args = (float(Cst), si, A, chr(int(tA)), sj, B, chr(int(tB)), sk, int(pol), int(zn))
        assert _cases.same_coo(orc.multiply_mm(*args), want), f"mm case {s}"
        if A.n * B.n < 5000:
            assert _cases.same_coo(orc.multiply_mm(*args, pairs=True), want), f"mm(pairs) case {s}"


def test_multiply_mv_fixtures(orc):
    p = _golden.pack("multiply_mv_cases")
    for s in range(int(p["count"])):
        si, A, sj, V, want = (_golden.get_coo(p, f"v{s}_{x}") for x in ("si", "A", "sj", "V", "out"))
        Cst, tA, pol, zn = p[f"v{s}_args"]
        got = orc.multiply_mv(float(Cst), si, A, chr(int(tA)), sj, V, int(pol), int(zn))
        assert _cases.same_coo(got, want), f"mv case {s}"


def test_config2_family_known_answer(orc):
    z = np.load(_golden.os.path.join(_golden.HERE, "golden", "config2_small.npz"))
    a = orc.gen_dup_coo(0x5EED0002, 0, 300000, 210000, 12, 1024)
    r = orc.consolidate(a, (0, 1))
    assert r.n == int(z["nnz"])
    w = np.arange(1, r.n + 1, dtype=np.uint64)
    assert [int((w * x.astype(np.uint64)).sum()) for x in r.idx] == [int(c) for c in z["chk"]]
    assert np.array_equal(r.idx[0][::64], z["idx0"]) and np.array_equal(r.val[::64], z["val"])


# ---- transpose / to_dense / to_sparse (algorithm.hpp:46-57, accum.hpp:110-140, algorithm.hpp:433-440)
def test_dense_ops_fixtures(orc):
    p = _golden.pack("dense_ops_cases")
    for s in range(int(p["count"])):
        a = _golden.get_coo(p, f"d{s}_in")
        args = [int(x) for x in p[f"d{s}_args"]]
        perm, policy = tuple(args[:-1]), args[-1]
        assert _cases.same_coo(orc.transpose(a, perm), _golden.get_coo(p, f"d{s}_T")), s
        dense = orc.to_dense(a, policy)
        want = p[f"d{s}_dense"]
        assert dense.shape == want.shape and np.array_equal(dense.view(np.uint64), want.view(np.uint64)), s  # bit for bit, NaN and -0 included
        assert _cases.same_coo(orc.to_sparse(dense), _golden.get_coo(p, f"d{s}_sparse")), s
