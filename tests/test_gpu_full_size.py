"""-m gpu: BASELINE configs 2-5 at their FULL sizes.  Two kinds of checks on every config:
  * size-independent properties evaluated on the device: the reference's known answer for config 2 (genuine
    reference, row-major consolidate of the 2*10^8-entry generator => nnz 139,999,960, first entry (0, 1022331), sum of
    values 200003308.953630), strict sortedness, idempotence; for config 5 the closed-form counts of the pentadiagonal
    product, row-major sortedness, exact linearity in a power-of-two scale; for config 3 structural symmetry;
  * a deterministic 1/64 ROW SAMPLE of the full-size GPU result compared with the CPU oracle bit for bit -- index structure,
    order AND values (SURVEY App. C parity rule).  Rows are independent in both operations (consolidate: an output row holds
    exactly the input entries of that row, in insertion order; multiply: multiply_sparse.hpp:192-246 carries no state from
    one row of A to the next), so the oracle run on the sampled rows of the SAME device-generated input is the sampled
    rows of the reference's full answer."""
import numpy as np
import pytest

from _gpu import DevView, dev_to_numpy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import spsparse_b200 as sp
    with sp.Context(0) as c:
        yield c


def views(a):
    import torch
    (p0, p1), pv = a.device_ptrs()
    n = a.size()
    return (torch.as_tensor(DevView(p0, n, "<i4"), device="cuda"), torch.as_tensor(DevView(p1, n, "<i4"), device="cuda"),
            torch.as_tensor(DevView(pv, n, "<f8"), device="cuda"))


def sample_mask(rows, block_bits=0):
    """deterministic 1/64 row sample: rows (or blocks of 2^block_bits consecutive rows) whose number is a multiple of 64"""
    return ((rows >> block_bits) & 63) == 0


def hashed_sample_mask(rows):
    """unbiased deterministic 1/64 row sample for power-law matrices (rows that are multiples of 64 are R-MAT's heavy rows:
    they hold a fifth of the products): top 6 bits of a 32-bit multiplicative hash of the row number.  Works on numpy
    arrays and torch tensors alike."""
    h = (rows.astype(np.int64) if isinstance(rows, np.ndarray) else rows.long()) * 0x9E3779B1
    return ((h >> 26) & 63) == 0


def host(*ts):
    return [t.cpu().numpy() for t in ts]


def same_bits(gi, gk, gv, want):
    return (want.n == len(gv) and np.array_equal(gi, want.idx[0]) and np.array_equal(gk, want.idx[1])
            and np.array_equal(gv.view(np.uint64), want.val.view(np.uint64)))


def strictly_ascending(i, k, bits):
    key = (i.long() << bits) | k.long()
    return bool((key[1:] > key[:-1]).all().item())


def test_config2_full_size(ctx):
    import spsparse_b200 as sp
    import torch
    n = 200_000_000
    A = sp.gen_dup_coo(ctx, 0x5EED0002, 0, n, 140_000_000, 24, 0)
    R, st = sp.consolidate(ctx, A, (0, 1), stats=True)
    i, k, v = views(R)
    assert R.size() == 139_999_960 and st.n_in == n and st.n_kept == n                  # the genuine reference's count
    assert (int(i[0].item()), int(k[0].item())) == (0, 1022331)                          # ... and first entry
    assert abs(float(v.sum().item()) - 200003308.953630) <= 1e-12 * 200003308.953630 * 50  # fp64 sum of 1.4e8 terms, different tree
    assert strictly_ascending(i, k, 24)                                                  # sorted and unique
    _, _, av = views(A)
    assert abs(float(v.sum().item()) - float(av.sum().item())) <= 1e-9 * float(av.sum().item())  # nothing lost
    R2 = sp.consolidate(ctx, R, (0, 1))                                                  # idempotent, bit for bit
    i2, k2, v2 = views(R2)
    assert R2.size() == R.size() and bool(torch.equal(i2, i)) and bool(torch.equal(k2, k)) and bool(torch.equal(v2, v))
    del i, k, v, i2, k2, v2, av
    for x in (A, R, R2):
        x.free()


@pytest.mark.parametrize("so", [(0, 1), (1, 0)])
def test_config2_row_sample_against_oracle(ctx, orc, so):
    """Config 2 at 2*10^8 entries with input zeros (every 1024th value, SURVEY App. C correctness variant): every 64th
    row (column for the column-major order) of the result against the oracle's consolidate of the input entries of
    those rows -- the default kernels at full size (row-digit passes, in-row column sort, reduce), values compared too."""
    import spsparse_b200 as sp
    from oracle import oracle as O
    n = 200_000_000
    A = sp.gen_dup_coo(ctx, 0x5EED0002, 0, n, 140_000_000, 24, 1024)
    R, st = sp.consolidate(ctx, A, so, stats=True)
    lead = so[0]
    ai = views(A)
    pick = sample_mask(ai[lead])
    a0, a1, av = host(ai[0][pick], ai[1][pick], ai[2][pick])
    ri = views(R)
    pick = sample_mask(ri[lead])
    g0, g1, gv = host(ri[0][pick], ri[1][pick], ri[2][pick])
    del ai, ri, pick
    assert st.n_kept < n and st.n_out == R.size()
    want = orc.consolidate(O.Coo((1 << 24, 1 << 24), [a0, a1], av), so)
    assert len(gv) > 2_000_000 and same_bits(g0, g1, gv, want)
    A.free(); R.free()


def test_config5_full_size(ctx, orc):
    import spsparse_b200 as sp
    import torch
    m = 100_000_000
    A, B, w = sp.gen_banded(ctx, 0x5EED0005, m, 0, m), sp.gen_banded(ctx, 0x5EED0015, m, 0, m), sp.gen_vector(ctx, 0x5EED0025, m)
    C1, st = sp.multiply(ctx, 1.0, None, A, ".", w, B, ".", None, stats=True)
    # pentadiagonal x pentadiagonal: 9 diagonals, 25 products per interior row
    assert st.products == 25 * m - 50 and C1.size() == 9 * m - 20 and st.rows_merge == m and st.rows_esc == 0
    i, k, v = views(C1)
    assert strictly_ascending(i, k, 27)                                                  # row-major, unique
    assert bool(((k - i).abs() <= 4).all().item())                                       # inside the 9-diagonal band
    assert bool((v > 0).all().item())                                                    # positive inputs: nothing cancels
    # linearity in a power of two is exact in floating point: C(4.0) == 4 * C(1.0) bit for bit
    C4 = sp.multiply(ctx, 4.0, None, A, ".", w, B, ".", None)
    i4, k4, v4 = views(C4)
    assert bool(torch.equal(i4, i)) and bool(torch.equal(k4, k)) and bool(torch.equal(v4, v * 4.0))
    del i4, k4, v4
    C4.free()
    # 1/64 sample in blocks of 1024 consecutive rows, bit for bit against the oracle (structure, order, values).  The sampled
    # rows of A need rows j-2..j+2 of B and the same entries of w: the oracle gets exactly those (a j absent from a scale
    # vector is excluded by the reference, so nothing else may be left out)
    from oracle import oracle as O
    pick = sample_mask(i, 10)
    gi, gk, gv = host(i[pick], k[pick], v[pick])
    del i, k, v, pick
    ai = views(A)
    pick = sample_mask(ai[0], 10)
    a0, a1, av = host(ai[0][pick], ai[1][pick], ai[2][pick])
    del ai, pick
    bi = views(B)
    need = torch.zeros(m, dtype=torch.bool, device="cuda")
    need[torch.as_tensor(np.unique(a1).astype(np.int64), device="cuda")] = True   # the inner indices the sampled rows hold
    pick = need[bi[0].long()]
    b0, b1, bv = host(bi[0][pick], bi[1][pick], bi[2][pick])
    del bi, pick
    (wp,), wv = w.device_ptrs()
    wi = torch.as_tensor(DevView(wp, m, "<i4"), device="cuda")
    wval = torch.as_tensor(DevView(wv, m, "<f8"), device="cuda")
    w0, w1 = host(wi[need], wval[need])
    del wi, wval, need
    want = orc.multiply_mm(1.0, None, O.Coo((m, m), [a0, a1], av), ".", O.Coo((m,), [w0], w1, (0,)),
                           O.Coo((m, m), [b0, b1], bv), ".", None)
    assert len(gv) > 14_000_000 and same_bits(gi, gk, gv, want)
    for x in (A, B, w, C1):
        x.free()


def test_config3_full_size(ctx, orc):
    """Regridding A*diag(s)*A^T, 10^7 x 10^6: the product is structurally symmetric -- transposing it on the device and
    consolidating gives back the same index structure; values agree to rounding (the two sides multiply in a
    different order)."""
    import spsparse_b200 as sp
    import torch
    A = sp.gen_regrid(ctx, 0x5EED0003, 3200, 3125, 1000, 1000)
    s = sp.gen_vector(ctx, 0x5EED0013, 1_000_000)
    C1, st = sp.multiply(ctx, 1.0, None, A, ".", s, A, "T", None, stats=True)
    assert st.rows_a == 10_000_000 and st.rows_merge == 10_000_000 and C1.size() == st.nnz_c
    i, k, v = views(C1)
    assert strictly_ascending(i, k, 24)
    Ct = sp.transpose(ctx, C1, (1, 0))
    Cs = sp.consolidate(ctx, Ct, (0, 1))
    Ct.free()
    it, kt, vt = views(Cs)
    assert Cs.size() == C1.size() and bool(torch.equal(it, i)) and bool(torch.equal(kt, k))
    assert bool(((vt - v).abs() <= 1e-12 * v.abs()).all().item())
    del it, kt, vt
    Cs.free()
    # every 64th row, bit for bit against the oracle: the sampled rows of A as the left operand, all of A (transposed by
    # the oracle itself, duplicates from the clamped last grid row/column included) as the right one
    from oracle import oracle as O
    pick = sample_mask(i)
    gi, gk, gv = host(i[pick], k[pick], v[pick])
    del i, k, v, pick
    ai = views(A)
    a0, a1, av = host(*ai)
    del ai
    (sp0,), spv = s.device_ptrs()
    s0 = dev_to_numpy(sp0, 1_000_000, "<i4")
    s1 = dev_to_numpy(spv, 1_000_000, "<f8")
    rows = (a0 & 63) == 0
    shp = (10_000_000, 1_000_000)
    want = orc.multiply_mm(1.0, None, O.Coo(shp, [a0[rows], a1[rows]], av[rows]), ".", O.Coo((1_000_000,), [s0], s1, (0,)),
                           O.Coo(shp, [a0, a1], av), "T", None)
    assert len(gv) > 10_000_000 and same_bits(gi, gk, gv, want)
    for x in (A, s, C1):
        x.free()


def test_config4_row_sample_against_oracle(ctx):
    """R-MAT scale 20 A*A (1.66e9 products, 1.2e9 outputs): every 64th row of the GPU result, bit for bit -- structure
    and values -- against the CPU oracle's product of those rows of A with A (SURVEY App. C parity rule for config 4)."""
    import spsparse_b200 as sp
    import torch
    from oracle import oracle as O
    from spsparse_b200 import gen
    sc = 20
    dA = sp.gen_rmat(ctx, 0x5EED0004, sc, 4 << sc)
    C1, st = sp.multiply(ctx, 1.0, None, dA, ".", None, dA, ".", None, stats=True)
    assert st.rows_hash > 100_000 and st.rows_merge > 100_000
    i, k, v = views(C1)
    assert strictly_ascending(i, k, sc)
    pick = (i % 64) == 0
    gi, gk, gv = i[pick].cpu().numpy(), k[pick].cpu().numpy(), v[pick].cpu().numpy()
    del i, k, v, pick
    shp, idx, val = gen.rmat(0x5EED0004, sc, 4 << sc)
    rows = (idx[0] % 64) == 0
    orc = O.port()
    want, wst = orc.multiply_mm(1.0, None, O.Coo(shp, [idx[0][rows], idx[1][rows]], val[rows]), ".", None,
                                O.Coo(shp, idx, val), ".", None, want_stats=True)
    assert want.n == len(gv) and np.array_equal(gi, want.idx[0]) and np.array_equal(gk, want.idx[1])
    assert np.array_equal(gv.view(np.uint64), want.val.view(np.uint64))
    dA.free(); C1.free()


def test_config4_as_named_scale24_row_panels(ctx, orc):
    """BASELINE config 4 as named: R-MAT 2^24 rows, edge factor 4 (67 M raw edges, duplicates kept), A*A.  The product has
    tens of billions of outputs -- more than any VectorCooArray can hold (algorithm.hpp:419) -- so it is formed in row
    panels (spb_mm_plan_*).  Every panel: row-major sorted and unique, rows inside the panel's range,
    counts equal to the symbolic-only call's.  A deterministic, unbiased 1/64 sample of the rows of the whole product, bit for bit -- structure, order and
    values -- against the CPU oracle (SURVEY App. C parity rule for config 4)."""
    import spsparse_b200 as sp
    import torch
    from oracle import oracle as O
    sc = 24
    dA = sp.gen_rmat(ctx, 0x5EED0004, sc, 4 << sc)
    a0, a1, av = host(*views(dA))
    plan = sp.MultiplyPlan(ctx, 1.0, None, dA, ".", None, dA, ".", None, max_products_per_panel=1 << 30)
    dA.free()
    assert plan.shape == (1 << sc, 1 << sc) and plan.n_panels >= 8
    gi, gk, gv = [], [], []
    F = nnz = 0
    last = -1
    for p in range(plan.n_panels):
        first_row, last_row, f = plan.info(p)
        C1, st = plan.panel(p, stats=True)
        assert st.products == f and C1.size() == st.nnz_c and first_row > last
        i, k, v = views(C1)
        assert strictly_ascending(i, k, sc)
        assert int(i[0].item()) >= first_row and int(i[-1].item()) <= last_row
        pick = hashed_sample_mask(i)
        for dst, t in zip((gi, gk, gv), host(i[pick], k[pick], v[pick])):
            dst.append(t)
        del i, k, v, pick
        C1.free()
        F += st.products; nnz += st.nnz_c; last = last_row
    assert F == plan.products
    plan.free()
    gi, gk, gv = np.concatenate(gi), np.concatenate(gk), np.concatenate(gv)
    rows = hashed_sample_mask(a0)
    shp = (1 << sc, 1 << sc)
    want = orc.multiply_mm(1.0, None, O.Coo(shp, [a0[rows], a1[rows]], av[rows]), ".", None, O.Coo(shp, [a0, a1], av), ".", None)
    print(f"config 4 at scale {sc}: F = {F}, nnz(C) = {nnz}, sampled outputs = {len(gv)}")
    assert len(gv) > nnz // 100 and same_bits(gi, gk, gv, want)
