"""-m gpu parity tests for consolidate / dim_beginnings: the CUDA path (through the C ABI) against
the reference-generated fixtures, the CPU oracle on fresh seeded inputs, and size-independent
properties at larger sizes.  Index structure must be bit-exact; values are bit-exact too because
the duplicate fold keeps the reference's left-to-right order (tolerance stated where it is not)."""
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


def gpu_consolidate(ctx, a, so, pol=O.ADD, zn=False):
    import spsparse_b200 as sp
    from _gpu import up, down
    A = up(ctx, a)
    R = sp.consolidate(ctx, A, so, pol, zn)
    out = down(R)
    db = R.dim_beginnings() if a.rank == 2 else None
    A.free(); R.free()
    return out, db


# tests/test_array.cpp:135-168
def test_known_answer(ctx):
    a = O.Coo((2, 4), [[1, 1, 0, 0, 1], [3, 2, 3, 1, 2]], [5., 3., 17., 14., 15.])
    r, db = gpu_consolidate(ctx, a, (0, 1))
    assert r.idx[0].tolist() == [0, 0, 1, 1] and r.idx[1].tolist() == [1, 3, 2, 3]
    assert r.val.tolist() == [14., 17., 18., 5.] and db.tolist() == [0, 2, 4]
    assert r.sort_order == (0, 1)
    r, db = gpu_consolidate(ctx, a, (1, 0))
    assert r.idx[0].tolist() == [0, 1, 0, 1] and r.idx[1].tolist() == [1, 2, 3, 3]
    assert r.val.tolist() == [14., 18., 17., 5.] and db.tolist() == [0, 1, 2, 4]


def test_fixtures_from_the_reference(ctx):
    p = _golden.pack("consolidate_cases")
    for s in range(int(p["count"])):
        a, want = _golden.get_coo(p, f"c{s}_in"), _golden.get_coo(p, f"c{s}_out")
        pol, zn, *so = (int(x) for x in p[f"c{s}_args"])
        got, db = gpu_consolidate(ctx, a, tuple(so), pol, zn)
        assert _cases.same_coo(got, want), f"consolidate case {s}"
        if a.rank == 2:
            assert np.array_equal(db, p[f"c{s}_db"]), f"dim_beginnings case {s}"


def test_against_oracle_fresh_seeds(ctx, orc):
    for s in range(1000, 1060):
        c = _cases.consolidate_case(s)
        a = O.Coo(tuple(c["shape"]), c["idx"], c["val"])
        args = (tuple(c["sort_order"]), c["policy"], c["zero_nan"])
        got, db = gpu_consolidate(ctx, a, *args)
        want = orc.consolidate(a, *args)
        assert _cases.same_coo(got, want), s
        if a.rank == 2:
            assert np.array_equal(db, orc.dim_beginnings(want)), s


def test_edge_cases(ctx, orc):
    import spsparse_b200 as sp
    # empty input: empty output, still flagged sorted (algorithm.hpp:263,318)
    r, db = gpu_consolidate(ctx, O.Coo((4, 4), [[], []], []), (0, 1))
    assert r.n == 0 and r.sort_order == (0, 1) and len(db) == 0
    # everything dropped
    r, _ = gpu_consolidate(ctx, O.Coo((4, 4), [[1, 2], [1, 2]], [0.0, -0.0]), (0, 1))
    assert r.n == 0
    r, _ = gpu_consolidate(ctx, O.Coo((4, 4), [[1, 2], [1, 2]], [np.nan, 0.0]), (0, 1), O.ADD, True)
    assert r.n == 0
    # one long run of duplicates (exercises the deferred long-run path); integer values => exact
    n = 20000
    a = O.Coo((8, 8), [np.full(n, 3), np.full(n, 5)], np.arange(1, n + 1, dtype=np.float64))
    for pol, want in ((O.ADD, n * (n + 1) / 2), (O.LEAVE_ALONE, 1.0), (O.REPLACE, float(n))):
        r, _ = gpu_consolidate(ctx, a, (0, 1), pol)
        assert r.n == 1 and r.val[0] == want
    # long run of reals: the deferred runs are folded left to right as well (one warp per run) => bit-identical
    rng = np.random.default_rng(3)
    a = O.Coo((8, 8), [np.full(n, 3), np.full(n, 5)], 0.5 + rng.random(n))
    r, _ = gpu_consolidate(ctx, a, (0, 1))
    want = orc.consolidate(a, (0, 1))
    assert r.val[0] == want.val[0]
    # maximum extents: 2^31 x 2^31 shape, indices at both ends
    big = (1 << 31) - 1
    a = O.Coo((1 << 31, 1 << 31), [[big, 0, big, 0], [big, big, big, 0]], [1., 2., 3., 4.])
    r, db = gpu_consolidate(ctx, a, (0, 1))
    assert r.idx[0].tolist() == [0, 0, big] and r.idx[1].tolist() == [0, big, big] and r.val.tolist() == [4., 2., 4.]
    # out-of-bounds index is an error, like VectorCooArray::add (VectorCooArray.hpp:245-262)
    A = sp.CooArray.from_host(ctx, (4, 4), [[1, 4], [1, 1]], [1., 1.])
    with pytest.raises(sp.SpbError):
        sp.consolidate(ctx, A, (0, 1))
    # dim_beginnings on an unsorted array is an error (algorithm.hpp:82-84)
    with pytest.raises(sp.SpbError) as e:
        A.dim_beginnings()
    assert e.value.code == 4
    A.free()


@pytest.mark.parametrize("so", [(0, 1), (1, 0)])
def test_duplicate_runs_of_every_length(ctx, orc, so):
    """Runs of 1 .. 6000 duplicates of real values (and zeros / NaNs to drop), scrambled: runs that end inside a thread's
    block, cross threads, cross tiles, and exceed the deferral threshold (RK_LONG_RUN = 256, finished by k_long_runs) --
    every sum bit-identical to the reference's left-to-right fold, for every policy."""
    rng = np.random.default_rng(41)
    lens = np.concatenate([rng.integers(1, 12, 3000), rng.integers(200, 330, 300), rng.integers(1000, 6000, 40), [257, 256, 255, 264, 2048, 2304]])
    keys = rng.permutation(len(lens))
    i = np.repeat(keys // 700, lens)
    k = np.repeat(keys % 700 * 1000, lens)
    o = rng.permutation(len(i))
    a = O.Coo((200, 1 << 20), [i[o], k[o]], _cases._values(rng, len(i), "mixed"))
    for pol in _cases.POLICIES:
        for zn in (0, 1):
            got, db = gpu_consolidate(ctx, a, so, pol, zn)
            want = orc.consolidate(a, so, pol, zn)
            assert _cases.same_coo(got, want), (pol, zn)


def test_default_path_large_against_oracle(ctx, orc):
    """2^23 + 12345 entries in a 10^8 x 10^8 shape: the size from which the DEFAULT path is row-digit passes with 9-bit
    digits (k_radix_pass9) + in-row column sort + reduce -- compared with the oracle in full, bit for bit: dropped zeros,
    zero_nan, duplicates, hub rows of thousands of entries (longer than the in-row sort handles: re-sorted by full key)."""
    import spsparse_b200 as sp
    from _gpu import up, down
    rng = np.random.default_rng(2024)
    n = (1 << 23) + 12345
    m = 100_000_000
    i = rng.integers(0, m, n)
    k = rng.integers(0, m, n)
    dup = rng.random(n) < 0.3                       # ~30 % duplicates of an earlier tuple
    src = rng.integers(0, n, n)
    i[dup], k[dup] = i[src[dup]], k[src[dup]]
    for h, row in enumerate((5, m - 1, 77_777_777)):  # hub rows
        sl = slice(100_000 * h, 100_000 * h + 3000 + 2000 * h)
        i[sl] = row
        k[sl] = rng.integers(0, 5000, sl.stop - sl.start)
    a = O.Coo((m, m), [i, k], _cases._values(rng, n, "mixed"))
    for so, pol, zn in (((0, 1), O.ADD, 1), ((1, 0), O.REPLACE, 0)):
        A = up(ctx, a)
        R, st = sp.consolidate(ctx, A, so, pol, zn, stats=True)
        got, db = down(R), R.dim_beginnings()
        A.free(); R.free()
        assert st.digit_bits == 9 and st.passes == 3   # the default took the path this test is about
        want = orc.consolidate(a, so, pol, zn)
        assert _cases.same_coo(got, want), (so, pol, zn)
        assert np.array_equal(db, orc.dim_beginnings(want))


def test_more_than_2_30_entries(ctx):
    """2^30 + 2^20 entries (the reference's cap is 2^31 - 1, algorithm.hpp:419-421; round 1 stopped at 2^30): properties
    on the device -- count, strict order, value sum, and the first 2^16 outputs against numpy."""
    import spsparse_b200 as sp
    import torch
    from _gpu import DevView
    n = (1 << 30) + (1 << 20)
    A = sp.gen_dup_coo(ctx, 0x5EED0002, 0, n, int(0.7 * n), 24, 0)
    R, st = sp.consolidate(ctx, A, (0, 1), stats=True)
    (p0, p1), pv = R.device_ptrs()
    m = R.size()
    assert st.n_in == n and st.n_kept == n and 0.69 * n < m <= 0.7 * n
    chunk = 1 << 28
    prev = -1
    total = 0.0
    for c0 in range(0, m, chunk):                   # in pieces: the int64 keys of 7.5e8 entries would not need to fit at once
        c1 = min(m, c0 + chunk)
        ti = torch.as_tensor(DevView(p0 + 4 * c0, c1 - c0, "<i4"), device="cuda").long()
        tk = torch.as_tensor(DevView(p1 + 4 * c0, c1 - c0, "<i4"), device="cuda").long()
        key = (ti << 24) | tk
        assert bool((key[1:] > key[:-1]).all().item()) and int(key[0].item()) > prev
        prev = int(key[-1].item())
        total += float(torch.as_tensor(DevView(pv + 8 * c0, c1 - c0, "<f8"), device="cuda").sum().item())
        del ti, tk, key
    (_, _), av = A.device_ptrs()
    asum = float(torch.as_tensor(DevView(av, n, "<f8"), device="cuda").sum().item())
    assert abs(total - asum) <= 1e-9 * asum
    A.free(); R.free()


def test_config2_family_known_answer(ctx):
    """Reduced-size member of BASELINE config 2; known answer from the genuine reference."""
    import spsparse_b200 as sp
    z = np.load(_golden.os.path.join(_golden.HERE, "golden", "config2_small.npz"))
    A = sp.gen_dup_coo(ctx, 0x5EED0002, 0, 300000, 210000, 12, 1024)
    R = sp.consolidate(ctx, A, (0, 1))
    idx, val = R.to_host()
    assert len(val) == int(z["nnz"])
    w = np.arange(1, len(val) + 1, dtype=np.uint64)
    assert [int((w * x.astype(np.uint64)).sum()) for x in idx] == [int(c) for c in z["chk"]]
    assert np.array_equal(idx[0][::64], z["idx0"]) and np.array_equal(val[::64], z["val"])
    A.free(); R.free()


def test_device_generator_matches_cpu(ctx, orc):
    import spsparse_b200 as sp
    from spsparse_b200 import gen
    A = sp.gen_dup_coo(ctx, 77, 1000, 50000, 30000, 10, 64)
    idx, val = A.to_host()
    want = orc.gen_dup_coo(77, 1000, 50000, 30000, 10, 64)
    npv = gen.dup_coo(77, 1000, 50000, 30000, 10, 64)
    for k in range(2):
        assert np.array_equal(idx[k], want.idx[k]) and np.array_equal(idx[k], npv[1][k])
    assert np.array_equal(val, want.val) and np.array_equal(val, npv[2])
    A.free()
    for name, dev, host in (("banded", sp.gen_banded(ctx, 5, 1000, 100, 700), gen.banded(5, 1000, 100, 700)),
                            ("regrid", sp.gen_regrid(ctx, 6, 32, 25, 10, 10), gen.regrid(6, 32, 25, 10, 10)),
                            ("rmat", sp.gen_rmat(ctx, 7, 10, 5000), gen.rmat(7, 10, 5000)),
                            ("vector", sp.gen_vector(ctx, 8, 333), gen.vector(8, 333))):
        idx, val = dev.to_host()
        assert dev.shape == tuple(host[0]), name
        for k in range(len(idx)):
            assert np.array_equal(idx[k], host[1][k]), name
        assert np.array_equal(val, host[2]), name
        dev.free()


def test_properties_at_scale(ctx):
    """2^24-entry member of config 2 (too big for the oracle in a unit test): sortedness, uniqueness,
    idempotence, count and value-sum conservation."""
    import spsparse_b200 as sp
    n, ub = 1 << 24, int(0.7 * (1 << 24))
    A = sp.gen_dup_coo(ctx, 0x5EED0002, 0, n, ub, 24, 0)
    R, st = sp.consolidate(ctx, A, (0, 1), stats=True)
    idx, val = R.to_host()
    _, aval = A.to_host()
    key = (idx[0].astype(np.int64) << 24) | idx[1]
    assert np.all(np.diff(key) > 0)                      # strictly ascending => sorted and unique
    assert st.n_in == n and st.n_kept == n and st.n_out == len(val)
    assert abs(val.sum() - aval.sum()) <= 1e-9 * aval.sum()
    assert ub - 64 <= len(val) <= ub                     # ~ub distinct tuples (a few random collisions)
    R2 = sp.consolidate(ctx, R, (0, 1))                   # idempotent
    idx2, val2 = R2.to_host()
    assert np.array_equal(idx2[0], idx[0]) and np.array_equal(idx2[1], idx[1]) and np.array_equal(val2, val)
    Rc = sp.consolidate(ctx, A, (1, 0))                   # the other order holds the same multiset
    idxc, valc = Rc.to_host()
    keyc = (idxc[1].astype(np.int64) << 24) | idxc[0]
    assert np.all(np.diff(keyc) > 0)
    o = np.argsort((idxc[0].astype(np.int64) << 24) | idxc[1], kind="stable")
    assert np.array_equal(idxc[0][o], idx[0]) and np.array_equal(idxc[1][o], idx[1]) and np.array_equal(valc[o], val)
    for x in (A, R, R2, Rc):
        x.free()


def test_medium_scale_against_oracle(ctx, orc):
    """4M entries of the config-2 family with input zeros, every policy and both orders, against the oracle."""
    import spsparse_b200 as sp
    from _gpu import down
    n = 4_000_000
    a = orc.gen_dup_coo(0x5EED0002, 0, n, int(0.7 * n), 16, 1024)
    A = sp.gen_dup_coo(ctx, 0x5EED0002, 0, n, int(0.7 * n), 16, 1024)
    for so, pol in (((0, 1), O.ADD), ((1, 0), O.ADD), ((0, 1), O.LEAVE_ALONE), ((1, 0), O.REPLACE)):
        R = sp.consolidate(ctx, A, so, pol)
        want = orc.consolidate(a, so, pol)
        assert _cases.same_coo(down(R), want), (so, pol)
        assert np.array_equal(R.dim_beginnings(), orc.dim_beginnings(want))
        assert np.array_equal(R.sorted_permutation(so), np.arange(R.size()))  # already sorted: identity
        R.free()
    perm = A.sorted_permutation((0, 1))
    assert np.array_equal(perm, orc.sorted_permutation(a, (0, 1)))
    A.free()


@pytest.mark.parametrize("mode,walk,fused,rwarp", [("1", "1", "0", None), ("1", "0", "0", None), ("0", "0", "0", None), ("1", "1", "1", None),
                                                   ("1", "1", "0", "0"), ("1", "0", "0", "2"), ("0", "0", "0", "2"),
                                                   ("1", "2", "0", None), ("1", "2", "0", "2")])
def test_row_passes_plus_in_row_column_sort(orc, monkeypatch, mode, walk, fused, rwarp):
    """The sort's second organisation -- radix passes over the row part of the key only, then every row ordered by
    column (k_segment_sort), rows longer than 64 entries re-sorted by their full key -- forced on (and off) for shapes
    with a wide column part: short rows, rows around the 64-entry limit, hub rows of thousands of entries next to
    short ones, all policies, zero_nan, both sort orders.  Same answers as the oracle, bit for bit."""
    import spsparse_b200 as sp
    from _gpu import up, down
    monkeypatch.setenv("SPB_SEGMENT_SORT", mode)
    monkeypatch.setenv("SPB_SEGMENT_WALK", walk)   # which in-row kernel: "1" neighbour walk, "0" row table, "2" shuffles (rows <= 5) + walk
    # fused = "1": the in-row sort runs inside the reduce pass (k_reduce_segsort); the cases with hub rows make it
    # give up and fall back to the separate kernels
    monkeypatch.setenv("SPB_FUSED_REDUCE", fused)
    # reduce pass: default = a warp per tile, tile offsets from the head counts of the in-row sort (no look-back) where one ran;
    # "0" = block per tile with look-back, "2" = warp per tile with look-back (also after full-key sorts)
    if rwarp is not None:
        monkeypatch.setenv("SPB_REDUCE_WARP", rwarp)
    rng = np.random.default_rng(11)
    cases = []
    for s, (shape, n, hubs) in enumerate([((300, 1 << 20), 5000, 0), ((3, 1 << 20), 4000, 0), ((2000, 1 << 18), 60000, 3),
                                          ((1 << 16, 1 << 16), 200000, 2), ((70, 1 << 17), 70 * 64, 0), ((1 << 20, 300), 30000, 1),
                                          ((5, 100000), 4099, 0), ((40000, 1 << 16), 1, 0),
                                          # no row over the limit (the fused kernel keeps its result): heavy duplicates,
                                          # very short rows, every row exactly at the 64-entry limit
                                          ((4000, 1 << 20), 50000, 0), ((1 << 17, 1 << 17), 300000, 0), ((1001, 1 << 16), 64007, 0)]):
        i = rng.integers(0, shape[0], n)
        k = rng.integers(0, min(shape[1], 5000 if s % 2 else shape[1]), n)     # odd cases: many duplicate tuples
        if s == 8:
            k = rng.integers(0, 6, n)                                           # runs of duplicates across tile ends
        if s == 10:
            i = np.concatenate([np.zeros(7, dtype=np.int64), np.repeat(np.arange(1, 1001), 64)])
            k = rng.integers(0, 40, n)
        for h in range(hubs):                                                   # hub rows: thousands of entries
            m = rng.integers(1000, 9000)
            i[h * 9000:h * 9000 + m] = 17 + 5 * h
        if s == 4:
            i = np.repeat(np.arange(70), 64)[:n] + 0                            # every row exactly 64 entries ...
            i[:130] = 3                                                          # ... one with 65+ (just over the limit)
        v = _cases._values(rng, n, ["pos", "int", "mixed"][s % 3])
        cases.append(O.Coo(shape, [i, k], v))
    with sp.Context(0) as c2:
        for s, a in enumerate(cases):
            for so in ((0, 1), (1, 0)):
                for pol in _cases.POLICIES:
                    zn = (s + pol) % 2
                    A = up(c2, a)
                    R, st = sp.consolidate(c2, A, so, pol, zn, stats=True)
                    got = down(R)
                    db = R.dim_beginnings()
                    A.free(); R.free()
                    want = orc.consolidate(a, so, pol, zn)
                    assert _cases.same_coo(got, want), (s, so, pol)
                    assert np.array_equal(db, orc.dim_beginnings(want)), (s, so, pol)
                if s in (0, 3, 8, 10):
                    # the stable argsort keeps every entry (KEEP_ALL through the same kernels: every entry is a run head)
                    A = up(c2, a)
                    perm = A.sorted_permutation(so)
                    A.free()
                    assert np.array_equal(perm, orc.sorted_permutation(a, so)), (s, so)


@pytest.mark.parametrize("seg", ["1", "0"])
def test_nine_bit_digit_passes(orc, monkeypatch, seg):
    """SPB_RADIX9=1: 9-bit digits (k_radix_pass9) wherever they cover the key -- or, with the in-row column sort, its row
    part -- in fewer passes than 8-bit ones: 17-bit and 27-bit row parts, 34- and 54-bit keys; all policies, zero_nan, both
    sort orders, duplicates.  Same answers as the oracle, bit for bit."""
    import spsparse_b200 as sp
    from _gpu import up, down
    monkeypatch.setenv("SPB_RADIX9", "1")
    monkeypatch.setenv("SPB_SEGMENT_SORT", seg)
    rng = np.random.default_rng(99)
    cases = []
    for s, (shape, n, kmax) in enumerate([((1 << 17, 1 << 17), 150000, None), ((100_000_000, 100_000_000), 120000, None),
                                          ((100_000_000, 100_000_000), 50000, 40), ((300, 1 << 20), 9000, None),
                                          ((1 << 26, 1 << 26), 4097, None), ((1 << 17, 1 << 17), 1, None)]):
        i = rng.integers(0, shape[0] if s != 2 else 3000, n)
        k = rng.integers(0, kmax or shape[1], n)
        cases.append(O.Coo(shape, [i, k], _cases._values(rng, n, ["pos", "int", "mixed"][s % 3])))
    with sp.Context(0) as c2:
        for s, a in enumerate(cases):
            for so in ((0, 1), (1, 0)):
                for pol in _cases.POLICIES:
                    zn = (s + pol) % 2
                    A = up(c2, a)
                    R, st = sp.consolidate(c2, A, so, pol, zn, stats=True)
                    got = down(R)
                    db = R.dim_beginnings()
                    A.free(); R.free()
                    want = orc.consolidate(a, so, pol, zn)
                    assert _cases.same_coo(got, want), (s, so, pol, st.passes)
                    assert np.array_equal(db, orc.dim_beginnings(want)), (s, so, pol)
