"""-m gpu parity tests for multiply (matrix*matrix and matrix*vector) through the C ABI."""
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


def gpu_mm(ctx, Cst, si, A, tA, sj, B, tB, sk, pol=O.ADD, zn=False):
    import spsparse_b200 as sp
    from _gpu import up, down
    hs = [up(ctx, x) for x in (si, A, sj, B, sk)]
    R = sp.multiply(ctx, Cst, hs[0], hs[1], tA, hs[2], hs[3], tB, hs[4], pol, zn)
    out = down(R)
    for h in hs + [R]:
        if h is not None:
            h.free()
    return out


def gpu_mv(ctx, Cst, si, A, tA, sj, V, pol=O.ADD, zn=False):
    import spsparse_b200 as sp
    from _gpu import up, down
    hs = [up(ctx, x) for x in (si, A, sj, V)]
    R = sp.multiply(ctx, Cst, hs[0], hs[1], tA, hs[2], hs[3], duplicate_policy=pol, zero_nan=zn)
    out = down(R)
    for h in hs + [R]:
        if h is not None:
            h.free()
    return out


# tests/test_multiply_sparse.cpp:41-79 (disabled in the reference, verified against it)
def test_scaled_known_answer(ctx):
    row = O.Coo((2, 10), [[0, 0, 0, 0, 1], [8, 4, 0, 3, 8]], [6., 4., 2., 3., 3.])
    scale = O.Coo((10,), [[0, 4, 8]], [2., 4., 4.], (0,))
    col = O.Coo((10, 1), [[0, 3, 8], [0, 0, 0]], [2., 3., 5.])
    eye = O.Coo((10,), [np.arange(10)], np.ones(10), (0,))
    r = gpu_mm(ctx, 1.0, eye, row, ".", scale, col, ".", eye)
    assert r.shape == (2, 1) and r.sort_order is None  # left in edit mode like the reference
    assert r.idx[0].tolist() == [0, 1] and r.idx[1].tolist() == [0, 0] and r.val.tolist() == [128., 60.]


# tests/test_multiply_sparse.cpp:84-136, :138-203 -- BASELINE config 1, seeds 1..999
def test_reference_random_tests(ctx):
    p = _golden.pack("reference_random_tests")
    eye = O.Coo((5,), [np.arange(5)], np.ones(5), (0,))
    for seed in range(1, 1000, 1 if _golden.os.environ.get("SPB_FULL_TESTS") else 3):
        A, B, Cg = (_golden.get_coo(p, f"mm{seed}_{x}") for x in "ABC")
        got = gpu_mm(ctx, 1.0, None, A, ".", eye, B, ".", None)
        assert _cases.same_coo(got, Cg), f"MM seed {seed}"
        A, V, Cg = (_golden.get_coo(p, f"mv{seed}_{x}") for x in "AVC")
        assert _cases.same_coo(gpu_mv(ctx, 1.0, None, A, ".", None, V), Cg), f"MV seed {seed}"


def test_mm_fixtures_from_the_reference(ctx):
    p = _golden.pack("multiply_mm_cases")
    for s in range(int(p["count"])):
        si, A, sj, B, sk, want = (_golden.get_coo(p, f"m{s}_{x}") for x in ("si", "A", "sj", "B", "sk", "out"))
        Cst, tA, tB, pol, zn = p[f"m{s}_args"]
        got = gpu_mm(ctx, float(Cst), si, A, chr(int(tA)), sj, B, chr(int(tB)), sk, int(pol), int(zn))
        assert _cases.same_coo(got, want), f"mm case {s}"


@pytest.mark.parametrize("variant,item_cap,win_cols", [(0, 0, 0), (1, 0, 0), (2, 0, 0), (0, 7, 0), (1, 64, 0), (2, 3, 0),
                                                       (2, 0, 32), (1, 5, 96), (0, 0, 1024)])
def test_mm_fixtures_through_the_hash_accumulator_bin(monkeypatch, variant, item_cap, win_cols):
    """Same fixtures with every row forced through the bitmap + shared-memory hash-accumulator kernels (both block
    shapes; with rows cut into many small work items so that the column windows are exercised): the sums are
    formed in ascending j, so every value is still bit-identical to the reference.  win_cols: bitmap of that many
    columns only (as for a matrix wider than the 1.5 M columns the real bitmap holds)."""
    import spsparse_b200 as sp
    _esc_env(monkeypatch, 0, 1 << 27, hash_min=0)
    monkeypatch.setenv("SPB_HASH_VARIANT", str(variant))
    if item_cap:
        monkeypatch.setenv("SPB_HASH_ITEM_CAP", str(item_cap))
    if win_cols:  # bitmap narrower than the matrix: rows are handled in several column windows
        monkeypatch.setenv("SPB_HASH_WIN_COLS", str(win_cols))
        # ... by the bitmap kernel: rows of a few thousand products would otherwise have their columns listed by the
        # shared-memory sort (k_hash_symbolic_small), which knows no windows -- the runs without win_cols go through that one
        monkeypatch.setenv("SPB_HASH_SMALL", "0" if variant != 1 else "1")
    else:
        monkeypatch.setenv("SPB_HASH_SMALL", "1" if variant == 2 else "0")
    p = _golden.pack("multiply_mm_cases")
    with sp.Context(0) as c2:
        for s in range(variant, int(p["count"]), 3 if item_cap == 0 else 5):
            si, A, sj, B, sk, want = (_golden.get_coo(p, f"m{s}_{x}") for x in ("si", "A", "sj", "B", "sk", "out"))
            Cst, tA, tB, pol, zn = p[f"m{s}_args"]
            got = gpu_mm(c2, float(Cst), si, A, chr(int(tA)), sj, B, chr(int(tB)), sk, int(pol), int(zn))
            assert _cases.same_coo(got, want), f"mm case {s}"


def test_mv_fixtures_from_the_reference(ctx):
    p = _golden.pack("multiply_mv_cases")
    for s in range(int(p["count"])):
        si, A, sj, V, want = (_golden.get_coo(p, f"v{s}_{x}") for x in ("si", "A", "sj", "V", "out"))
        Cst, tA, pol, zn = p[f"v{s}_args"]
        got = gpu_mv(ctx, float(Cst), si, A, chr(int(tA)), sj, V, int(pol), int(zn))
        assert _cases.same_coo(got, want), f"mv case {s}"


def test_mm_against_oracle_fresh_seeds(ctx, orc):
    for s in range(2000, 2080):
        c = _cases.mm_case(s, big=(s % 4 == 0))
        args = (c["C"], c["si"], c["A"], c["tA"], c["sj"], c["B"], c["tB"], c["sk"], c["policy"], c["zero_nan"])
        assert _cases.same_coo(gpu_mm(ctx, *args), orc.multiply_mm(*args)), s


def test_cancellation_in_the_register_merge_bin(orc):
    """Short rows whose terms cancel to exactly 0: the reference drops such outputs (multiply_sparse.hpp:238).  The symbolic
    pass of the register-merge bin counts distinct columns without reading a value, so the numeric pass leaves tombstones
    that the compaction closes -- whole rows vanishing, partial cancellation, NaN sums (kept), next to ordinary rows, also
    in row panels."""
    import spsparse_b200 as sp
    from _gpu import up, down
    rng = np.random.default_rng(404)
    m, nj, nk = 3000, 400, 5000
    # B: rows 2j and 2j+1 identical for j < 100 (so +a / -a on that pair cancels every shared output), the rest random
    bj, bk, bv = [], [], []
    for j in range(100):
        cols = np.sort(rng.choice(nk, 6, replace=False))
        vals = rng.integers(1, 5, 6).astype(np.float64)
        for jj in (2 * j, 2 * j + 1):
            bj += [jj] * 6; bk += cols.tolist(); bv += vals.tolist()
    nb = 3000
    bj += rng.integers(200, nj, nb).tolist(); bk += rng.integers(0, nk, nb).tolist(); bv += rng.integers(1, 5, nb).astype(float).tolist()
    B = O.Coo((nj, nk), [np.array(bj), np.array(bk)], np.array(bv))
    ai, aj, av = [], [], []
    for i in range(m):
        kind = i % 4
        if kind == 0:      # everything cancels: the row of C is empty
            p = int(rng.integers(0, 100)); ai += [i, i]; aj += [2 * p, 2 * p + 1]; av += [3.0, -3.0]
        elif kind == 1:    # a cancelling pair plus other terms: some outputs vanish, some survive
            p = int(rng.integers(0, 100)); q = int(rng.integers(200, nj))
            ai += [i, i, i]; aj += [2 * p, 2 * p + 1, q]; av += [2.0, -2.0, 1.5]
        elif kind == 2:    # no cancellation
            for q in rng.choice(np.arange(200, nj), 4, replace=False):
                ai.append(i); aj.append(int(q)); av.append(float(rng.integers(1, 4)))
        else:              # +inf - inf = NaN: kept (NaN != 0)
            p = int(rng.integers(0, 100)); ai += [i, i]; aj += [2 * p, 2 * p + 1]; av += [np.inf, -np.inf]
    A = O.Coo((m, nj), [np.array(ai), np.array(aj)], np.array(av))
    want, st = orc.multiply_mm(1.0, None, A, ".", None, B, ".", None, want_stats=True)
    with sp.Context(0) as c2:
        hs = [up(c2, x) for x in (A, B)]
        R, gst = sp.multiply(c2, 1.0, None, hs[0], ".", None, hs[1], ".", None, stats=True)
        got = down(R)
        assert gst.products == st["F"] and gst.rows_hash == 0 and gst.rows_esc == 0 and gst.rows_merge > 0
        assert gst.nnz_c == want.n and _cases.same_coo(got, want)
        assert np.isnan(got.val).sum() == np.isnan(want.val).sum() > 0
        for h in hs + [R]:
            h.free()
        got_p, n_panels = gpu_mm_panels(c2, 1.0, None, A, ".", None, B, ".", None, max_products=2000)
        assert n_panels > 3 and _cases.same_coo(got_p, want)


def test_errors_and_empties(ctx):
    import spsparse_b200 as sp
    A = O.Coo((2, 3), [[0], [0]], [1.])
    B = O.Coo((2, 2), [[0], [0]], [1.])
    with pytest.raises(sp.SpbError) as e:  # multiply_sparse.hpp:172-174
        gpu_mm(ctx, 1.0, None, A, ".", None, B, ".", None)
    assert e.value.code == 3
    B = O.Coo((3, 2), [[0], [0]], [1.])
    empty = O.Coo((2, 3), [[], []], [])
    assert gpu_mm(ctx, 0.0, None, A, ".", None, B, ".", None).n == 0        # C == 0  (:178)
    r = gpu_mm(ctx, 1.0, None, empty, ".", None, B, ".", None)                # empty A
    assert r.n == 0 and r.shape == (2, 2)
    assert gpu_mm(ctx, 1.0, O.Coo((2,), [[]], [], (0,)), A, ".", None, B, ".", None).n == 0  # empty scale
    allzero = O.Coo((2, 3), [[0, 1], [0, 2]], [0., 0.])                        # consolidates to nothing
    assert gpu_mm(ctx, 1.0, None, allzero, ".", None, B, ".", None).n == 0


def _esc_env(monkeypatch, merge_max, chunk, hash_min="off"):
    monkeypatch.setenv("SPB_MERGE_MAX_PRODUCTS", str(merge_max))
    monkeypatch.setenv("SPB_ESC_CHUNK", str(chunk))
    monkeypatch.setenv("SPB_HASH_MIN_PRODUCTS", str(hash_min))  # "off": long rows all go through expand-sort-compress


def test_expand_sort_compress_path(orc, monkeypatch):
    """Force every row through expand-sort-compress (and through several chunks): same answers."""
    import spsparse_b200 as sp
    _esc_env(monkeypatch, 0, 37)
    with sp.Context(0) as c2:
        for s in list(range(2100, 2130)) + [2400, 2404]:
            c = _cases.mm_case(s, big=(s % 4 == 0))
            args = (c["C"], c["si"], c["A"], c["tA"], c["sj"], c["B"], c["tB"], c["sk"], c["policy"], c["zero_nan"])
            assert _cases.same_coo(gpu_mm(c2, *args), orc.multiply_mm(*args)), s


def test_mixed_bins_medium(orc, monkeypatch):
    """A few thousand rows with a handful of heavy rows: merge rows and ESC rows interleave."""
    import spsparse_b200 as sp
    rng = np.random.default_rng(11)
    m, nj, nk = 3000, 2500, 2000
    n = 12000
    ai, aj = rng.integers(0, m, n), rng.integers(0, nj, n)
    heavy = rng.choice(m, 5, replace=False)
    ai = np.concatenate([ai, np.repeat(heavy, 300)]); aj = np.concatenate([aj, rng.integers(0, nj, 1500)])
    A = O.Coo((m, nj), [ai, aj], 0.5 + rng.random(len(ai)))
    nb = 15000
    B = O.Coo((nj, nk), [rng.integers(0, nj, nb), rng.integers(0, nk, nb)], 0.5 + rng.random(nb))
    sj = O.Coo((nj,), [np.arange(nj)], 0.5 + rng.random(nj), (0,))
    want, st = orc.multiply_mm(1.0, None, A, ".", sj, B, ".", None, want_stats=True)
    _esc_env(monkeypatch, 64, 5000)
    with sp.Context(0) as c2:
        from _gpu import up, down
        hs = [up(c2, x) for x in (A, sj, B)]
        R, gst = sp.multiply(c2, 1.0, None, hs[0], ".", hs[1], hs[2], ".", None, stats=True)
        got = down(R)
        assert gst.products == st["F"] and gst.nnz_a == st["nnzA"] and gst.nnz_b == st["nnzB"]
        assert gst.rows_esc > 0 and gst.rows_merge > 0
        assert _cases.same_coo(got, want)
        # transposed route: (A*diag(sj)*B)^T == B^T * diag(sj) * A^T ; same multiset of entries
        Rt = sp.multiply(c2, 1.0, None, hs[2], "T", hs[1], hs[0], "T", None)
        gt = down(Rt)
        o = np.lexsort((gt.idx[0], gt.idx[1]))
        assert np.array_equal(gt.idx[1][o], got.idx[0]) and np.array_equal(gt.idx[0][o], got.idx[1])
        assert np.allclose(gt.val[o], got.val, rtol=1e-12, atol=0)  # different summation order: 1e-12 relative
        for h in hs + [R, Rt]:
            h.free()


def test_regrid_and_banded_families_small(ctx, orc):
    """Reduced-size members of BASELINE configs 3 and 5 against the CPU oracle."""
    import spsparse_b200 as sp
    from spsparse_b200 import gen
    from _gpu import down
    # config 3: A (ny*nx x gy*gx), C = A diag(s) A^T
    shp, idx, val = gen.regrid(0x5EED0003, 64, 50, 20, 16)
    s = gen.vector(0x5EED0013, shp[1])
    A = O.Coo(shp, idx, val)
    S = O.Coo(s[0], s[1], s[2], (0,))
    want, st = orc.multiply_mm(1.0, None, A, ".", S, A, "T", None, want_stats=True)
    dA = sp.gen_regrid(ctx, 0x5EED0003, 64, 50, 20, 16)
    dS = sp.gen_vector(ctx, 0x5EED0013, shp[1])
    R, gst = sp.multiply(ctx, 1.0, None, dA, ".", dS, dA, "T", None, stats=True)
    assert _cases.same_coo(down(R), want)
    assert gst.products == st["F"] and gst.nnz_c == want.n
    for h in (dA, dS, R):
        h.free()
    # config 5: pentadiagonal A diag(w) B
    m = 5000
    a = gen.banded(0x5EED0005, m, 0, m); b = gen.banded(0x5EED0015, m, 0, m); w = gen.vector(0x5EED0025, m)
    want, st = orc.multiply_mm(1.0, None, O.Coo(*a), ".", O.Coo(w[0], w[1], w[2], (0,)), O.Coo(*b), ".", None, want_stats=True)
    dA, dB, dW = sp.gen_banded(ctx, 0x5EED0005, m, 0, m), sp.gen_banded(ctx, 0x5EED0015, m, 0, m), sp.gen_vector(ctx, 0x5EED0025, m)
    R, gst = sp.multiply(ctx, 1.0, None, dA, ".", dW, dB, ".", None, stats=True)
    assert _cases.same_coo(down(R), want)
    assert gst.products == st["F"]
    assert want.n == 9 * m - 20
    for h in (dA, dB, dW, R):
        h.free()


def test_rmat_family_small(ctx, orc):
    """Reduced-size member of BASELINE config 4 (skewed rows: both bins in use)."""
    import spsparse_b200 as sp
    from spsparse_b200 import gen
    from _gpu import down
    shp, idx, val = gen.rmat(0x5EED0004, 12, 4 << 12)
    A = O.Coo(shp, idx, val)
    want, st = orc.multiply_mm(1.0, None, A, ".", None, A, ".", None, want_stats=True)
    dA = sp.gen_rmat(ctx, 0x5EED0004, 12, 4 << 12)
    R, gst = sp.multiply(ctx, 1.0, None, dA, ".", None, dA, ".", None, stats=True)
    assert gst.products == st["F"] and gst.rows_hash > 0 and gst.rows_merge > 0
    assert _cases.same_coo(down(R), want)
    dA.free(); R.free()


def test_prepared_and_compressed_operands(ctx, orc):
    """multiply() == consolidate + multiply_prepared, with B as COO and with B in compressed (pointer) form
    -- the route the multi-GPU path takes."""
    import spsparse_b200 as sp
    from spsparse_b200 import gen
    from _gpu import down
    m = 3000
    dA, dB, dW = sp.gen_banded(ctx, 0x51, m, 0, m), sp.gen_banded(ctx, 0x52, m, 0, m), sp.gen_vector(ctx, 0x53, m)
    want = down(sp.multiply(ctx, 1.0, None, dA, ".", dW, dB, ".", None))
    Ac, Bc = sp.consolidate(ctx, dA, (0, 1)), sp.consolidate(ctx, dB, (0, 1))
    C1, st1 = sp.multiply_prepared(ctx, 1.0, None, Ac, 0, dW, Bc, 0, None)
    assert _cases.same_coo(down(C1), want)
    (p0, p1), pv = Bc.device_ptrs()
    ptr, ext = Bc.dense_ptr()
    assert ext == m
    Bcsr = sp.CooArray.wrap_csr(ctx, (m, m), 0, ptr, p1, pv, Bc.size())
    C2, st2 = sp.multiply_prepared(ctx, 1.0, None, Ac, 0, dW, Bcsr, 0, None)
    assert _cases.same_coo(down(C2), want) and st2.products == st1.products
    with pytest.raises(sp.SpbError):
        sp.consolidate(ctx, Bcsr, (0, 1))  # a compressed-form array is a B operand only
    # ranged pointer == the same stretch of the full one (what a rank of the row-partitioned multiply asks for)
    from _gpu import dev_to_numpy
    full = dev_to_numpy(ptr, m + 1, "<i4")
    for lo, hi in ((0, m), (17, 1203), (m - 5, m), (40, 40)):
        part = dev_to_numpy(Bc.dense_ptr_range(lo, hi), hi - lo + 1, "<i4")
        assert np.array_equal(part, full[lo:hi + 1]), (lo, hi)
    for h in (dA, dB, dW, Ac, C1, C2, Bcsr, Bc):
        h.free()


def test_medium_scale_against_oracle(ctx, orc):
    """Mid-size members of BASELINE configs 3, 4 and 5 (10^5..10^6 rows, millions of products): full
    bit-for-bit comparison with the CPU oracle's row-wise evaluation -- structure AND values."""
    import spsparse_b200 as sp
    from spsparse_b200 import gen
    from _gpu import down
    # config 5 family, 400k rows (2M entries per operand, ~10M products)
    m = 400_000
    a, b, w = gen.banded(0x5EED0005, m, 0, m), gen.banded(0x5EED0015, m, 0, m), gen.vector(0x5EED0025, m)
    want, st = orc.multiply_mm(1.0, None, O.Coo(*a), ".", O.Coo(w[0], w[1], w[2], (0,)), O.Coo(*b), ".", None, want_stats=True)
    dA, dB, dW = sp.gen_banded(ctx, 0x5EED0005, m, 0, m), sp.gen_banded(ctx, 0x5EED0015, m, 0, m), sp.gen_vector(ctx, 0x5EED0025, m)
    R, gst = sp.multiply(ctx, 1.0, None, dA, ".", dW, dB, ".", None, stats=True)
    assert gst.products == st["F"] and _cases.same_coo(down(R), want)
    for h in (dA, dB, dW, R):
        h.free()
    # config 3 family, 320x312 fine grid on 100x100 coarse grid (~16M products)
    shp, idx, val = gen.regrid(0x5EED0003, 320, 312, 100, 100)
    s = gen.vector(0x5EED0013, shp[1])
    want, st = orc.multiply_mm(1.0, None, O.Coo(shp, idx, val), ".", O.Coo(s[0], s[1], s[2], (0,)), O.Coo(shp, idx, val), "T", None, want_stats=True)
    dA, dS = sp.gen_regrid(ctx, 0x5EED0003, 320, 312, 100, 100), sp.gen_vector(ctx, 0x5EED0013, shp[1])
    R, gst = sp.multiply(ctx, 1.0, None, dA, ".", dS, dA, "T", None, stats=True)
    assert gst.products == st["F"] and _cases.same_coo(down(R), want)
    for h in (dA, dS, R):
        h.free()
    # config 4 family, R-MAT scale 15 (skewed: most products go through the hash-accumulator bin, the hub rows in
    # several column windows)
    shp, idx, val = gen.rmat(0x5EED0004, 15, 4 << 15)
    want, st = orc.multiply_mm(1.0, None, O.Coo(shp, idx, val), ".", None, O.Coo(shp, idx, val), ".", None, want_stats=True)
    dA = sp.gen_rmat(ctx, 0x5EED0004, 15, 4 << 15)
    R, gst = sp.multiply(ctx, 1.0, None, dA, ".", None, dA, ".", None, stats=True)
    assert gst.products == st["F"] and gst.products_hash > 0 and _cases.same_coo(down(R), want)
    dA.free(); R.free()


@pytest.mark.parametrize("variant,win_cols,walk", [(0, 0, None), (1, 0, None), (2, 0, None), (2, 4096, None), (2, 0, "0"), (2, 4096, "0")])
def test_hash_accumulator_bin(orc, monkeypatch, variant, win_cols, walk):
    """Long rows through the bitmap + hash-accumulator kernels: all three bins side by side, scale vectors, a row
    wider than one work item, and rows whose sums cancel exactly (they emit fewer entries than the symbolic
    bound -> gap compaction).  Bit-identical values throughout."""
    import spsparse_b200 as sp
    from _gpu import up, down
    rng = np.random.default_rng(21)
    m, nj, nk = 400, 300, 30000
    # ordinary sparse part
    n = 1200
    ai, aj = rng.integers(0, m, n), rng.integers(0, nj, n)
    av = rng.integers(1, 4, n).astype(np.float64)
    nb, nb2 = 20000, 40000  # inner indices 200..299 hold most of B: the heavy rows below use only those
    bj, bk = np.concatenate([rng.integers(0, nj, nb), rng.integers(200, nj, nb2)]), rng.integers(0, nk, nb + nb2)
    bv = rng.integers(1, 4, nb + nb2).astype(np.float64)
    # rows 0..3 of A: +1 on inner index 10, -1 on inner index 11; B rows 10 and 11 identical => exact cancellation
    canc = rng.choice(nk, 700, replace=False)
    keep_b = (bj != 10) & (bj != 11)
    bj, bk, bv = bj[keep_b], bk[keep_b], bv[keep_b]
    bj = np.concatenate([bj, np.full(700, 10), np.full(700, 11)]); bk = np.concatenate([bk, canc, canc])
    bv = np.concatenate([bv, np.full(1400, 2.0)])
    keep_a = ai > 3
    ai, aj, av = ai[keep_a], aj[keep_a], av[keep_a]
    ai = np.concatenate([ai, np.repeat(np.arange(4), 2), [2, 3]]); aj = np.concatenate([aj, np.tile([10, 11], 4), [50, 60]])
    av = np.concatenate([av, np.tile([1.0, -1.0], 4), [5.0, 7.0]])
    # a few heavy rows with real-valued data
    heavy = np.arange(100, 104)
    ai = np.concatenate([ai, np.repeat(heavy, 120)]); aj = np.concatenate([aj, rng.integers(200, nj, 480)])  # ~20k outputs each
    av = np.concatenate([av, 0.5 + rng.random(480)])
    A = O.Coo((m, nj), [ai, aj], av)
    B = O.Coo((nj, nk), [bj, bk], bv)
    si = O.Coo((m,), [np.arange(0, m, 2)], 1.0 + np.arange(0, m, 2) % 3, (0,))
    sk = O.Coo((nk,), [np.arange(1, nk)], np.where(np.arange(1, nk) % 7 == 0, 0.0, 2.0), (0,))
    for scales in ((None, None), (si, sk)):
        want, st = orc.multiply_mm(1.5, scales[0], A, ".", None, B, ".", scales[1], want_stats=True)
        _esc_env(monkeypatch, 256, 4000, hash_min=512)
        monkeypatch.setenv("SPB_HASH_VARIANT", str(variant))
        if win_cols:
            monkeypatch.setenv("SPB_HASH_WIN_COLS", str(win_cols))
        if walk is not None:   # "0": the bitmap kernel walks every 32-word group of every unit (the path before the summary bits)
            monkeypatch.setenv("SPB_HASH_SPARSE_WALK", walk)
        with sp.Context(0) as c2:
            hs = [up(c2, x) for x in (scales[0], A, B, scales[1])]
            R, gst = sp.multiply(c2, 1.5, hs[0], hs[1], ".", None, hs[2], ".", hs[3], stats=True)
            got = down(R)
            assert gst.products == st["F"] and gst.rows_hash >= 4 and gst.rows_merge > 0 and gst.rows_esc > 0
            assert _cases.same_coo(got, want), scales[0] is not None
            for h in hs + [R]:
                if h is not None:
                    h.free()


def test_hash_bin_wider_than_the_bitmap(ctx, orc):
    """4 M output columns (the shared-memory bitmap holds 1.5 M): long rows go through the hash bin in three column
    windows, their outputs cut into many work items; short rows through the merge.  Bit-identical to the oracle."""
    import spsparse_b200 as sp
    from _gpu import up, down
    rng = np.random.default_rng(77)
    m, nj, nk = 260, 1500, 4_000_000
    na = 12000
    ai, aj = rng.integers(0, 200, na), rng.integers(0, nj, na)          # rows 0..199: ~60 entries each
    ai = np.concatenate([ai, np.arange(200, 260)]); aj = np.concatenate([aj, rng.integers(0, nj, 60)])  # rows 200..259: one entry
    av = rng.standard_normal(len(ai))
    nb = 1_500_000
    bj, bk = rng.integers(0, nj, nb), rng.integers(0, nk, nb)
    bv = rng.standard_normal(nb)
    A, B = O.Coo((m, nj), [ai, aj], av), O.Coo((nj, nk), [bj, bk], bv)
    sk = O.Coo((nk,), [np.arange(0, nk, 3)], 1.0 + (np.arange(0, nk, 3) % 5), (0,))   # two thirds of the columns excluded
    for scalek in (None, sk):
        want, st = orc.multiply_mm(0.5, None, A, ".", None, B, ".", scalek, want_stats=True)
        hs = [up(ctx, x) for x in (A, B, scalek)]
        R, gst = sp.multiply(ctx, 0.5, None, hs[0], ".", None, hs[1], ".", hs[2], stats=True)
        assert gst.products == st["F"] and gst.rows_hash >= 190 and gst.rows_merge > 0, gst.asdict()
        assert _cases.same_coo(down(R), want), scalek is not None
        for h in hs + [R]:
            if h is not None:
                h.free()


# ---- multiply in row panels (spb_mm_plan_*): the panels concatenated ARE the product ---------------------------------------
def gpu_mm_panels(ctx, Cst, si, A, tA, sj, B, tB, sk, pol=O.ADD, zn=False, max_products=1 << 30):
    import spsparse_b200 as sp
    from _gpu import up, down
    hs = [up(ctx, x) for x in (si, A, sj, B, sk)]
    plan = sp.MultiplyPlan(ctx, Cst, hs[0], hs[1], tA, hs[2], hs[3], tB, hs[4], pol, zn, max_products)
    hs[1].free(); hs[3].free()                 # the plan holds its own consolidated operands
    parts, sym = [], []
    last = -1
    for p in range(plan.n_panels):
        first_row, last_row, f = plan.info(p)
        st0 = plan.symbolic(p)
        R, st = plan.panel(p, stats=True)
        c = down(R)
        R.free()
        assert st0.products == st.products == f and st0.rows_merge == st.rows_merge and st0.rows_hash == st.rows_hash
        assert st0.nnz_c >= st.nnz_c == c.n     # the symbolic count is exact unless hash-bin sums cancel to 0
        assert first_row > last and last_row >= first_row
        if c.n:
            assert first_row <= c.idx[0].min() and c.idx[0].max() <= last_row
        last = last_row
        parts.append(c)
        sym.append(st0)
    shape, total = plan.shape, plan.products
    plan.free()
    for h in (hs[0], hs[2], hs[4]):
        if h is not None:
            h.free()
    idx = [np.concatenate([c.idx[d] for c in parts]) if parts else np.empty(0, np.int32) for d in (0, 1)]
    val = np.concatenate([c.val for c in parts]) if parts else np.empty(0)
    assert sum(s.products for s in sym) == total
    return O.Coo(shape, idx, val, None), len(parts)


@pytest.mark.parametrize("max_products", [1, 7, 1 << 30])
def test_mm_fixtures_in_row_panels(ctx, max_products):
    p = _golden.pack("multiply_mm_cases")
    for s in range(0, int(p["count"]), 2):
        si, A, sj, B, sk, want = (_golden.get_coo(p, f"m{s}_{x}") for x in ("si", "A", "sj", "B", "sk", "out"))
        Cst, tA, tB, pol, zn = p[f"m{s}_args"]
        got, n_panels = gpu_mm_panels(ctx, float(Cst), si, A, chr(int(tA)), sj, B, chr(int(tB)), sk, int(pol), int(zn), max_products)
        assert _cases.same_coo(got, want), f"mm case {s} in {n_panels} panels"


@pytest.mark.parametrize("hash_min", [512, 0])
def test_rmat_in_row_panels_against_oracle(orc, monkeypatch, hash_min):
    """R-MAT scale 13 A*A (hub rows, all three bins) cut into ~40 panels: identical to the oracle's product, and the
    symbolic-only counts add up to the product's."""
    import spsparse_b200 as sp
    from spsparse_b200 import gen
    monkeypatch.setenv("SPB_HASH_MIN_PRODUCTS", str(hash_min))
    shp, idx, val = gen.rmat(0x5EED0004, 13, 4 << 13)
    a = O.Coo(shp, idx, val)
    want, wst = orc.multiply_mm(1.0, None, a, ".", None, a, ".", None, want_stats=True)
    with sp.Context(0) as c2:
        got, n_panels = gpu_mm_panels(c2, 1.0, None, a, ".", None, a, ".", None, max_products=wst["F"] // 40)
    assert 20 <= n_panels <= 41 and _cases.same_coo(got, want)


def test_scale_vector_must_be_ascending(ctx):
    """The reference joins scale vectors as sorted lists without repeats (xiter.hpp:146, 201); one that is not has no
    defined product there -- here it is an argument error, not a race between the repeated entries."""
    import spsparse_b200 as sp
    A = O.Coo((3, 4), [[0, 1, 2], [0, 1, 3]], [1., 2., 3.])
    B = O.Coo((4, 2), [[0, 1, 3], [0, 1, 1]], [1., 1., 1.])
    for bad in (O.Coo((4,), [[0, 1, 1, 3]], [1., 2., 5., 1.], (0,)), O.Coo((4,), [[0, 3, 1]], [1., 2., 5.], (0,))):
        with pytest.raises(sp.SpbError) as e:
            gpu_mm(ctx, 1.0, None, A, ".", bad, B, ".", None)
        assert e.value.code == 2
    V = O.Coo((4,), [[1, 1]], [1., 2.], (0,))      # flagged sorted by the caller, but it is not
    with pytest.raises(sp.SpbError):
        gpu_mv(ctx, 1.0, None, A, ".", None, V)


def test_pool_trim(ctx):
    import ctypes
    import spsparse_b200 as sp
    A = sp.gen_dup_coo(ctx, 1, 0, 1 << 20, 1 << 19, 12, 0)
    sp.consolidate(ctx, A, (0, 1)).free()
    A.free()
    freed = ctypes.c_uint64()
    assert ctx.lib.spb_ctx_trim(ctx.h, ctypes.byref(freed)) == 0 and freed.value >= (1 << 20) * 16
    assert ctx.lib.spb_ctx_trim(ctx.h, ctypes.byref(freed)) == 0 and freed.value == 0


def test_row_partitioned_multiply_on_two_gpus():
    """spb_rowpart_* (one process per GPU, shards of B fetched from the peers' memory): tests/multi/rowpart_check.py under
    torchrun on 2 GPUs -- skipped on a single-GPU box (bench.py --gpus N checks the result fingerprint at every N)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(root, "tests", "multi", "rowpart_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_row_partition_with_one_rank(ctx, orc):
    """n_ranks = 1: spb_rowpart_multiply is consolidate(B) + consolidate(A) + multiply, no peers -- same answer as the oracle."""
    import spsparse_b200 as sp
    from spsparse_b200.dist import RowPartition
    from _gpu import up, down
    for s in (2003, 2004, 2008):
        c = _cases.mm_case(s, big=True)
        if c["tA"] != "." or c["tB"] != ".":
            c = dict(c, A=orc.transpose(c["A"], (1, 0)) if c["tA"] == "T" else c["A"], B=orc.transpose(c["B"], (1, 0)) if c["tB"] == "T" else c["B"])
        m = c["B"].shape[0]
        rp = RowPartition(ctx, 0, 1, m, c["B"].n + 1)
        hs = [up(ctx, x) for x in (c["si"], c["A"], c["sj"], c["B"], c["sk"])]
        Cm, st = rp.multiply(c["C"], hs[0], hs[1], hs[2], hs[3], hs[4], c["policy"], c["zero_nan"])
        got = down(Cm)
        for h in hs + [Cm]:
            if h is not None:
                h.free()
        rp.close()
        want = orc.multiply_mm(c["C"], c["si"], c["A"], ".", c["sj"], c["B"], ".", c["sk"], c["policy"], c["zero_nan"])
        assert _cases.same_coo(got, want), s
