"""-m gpu: BASELINE configs 2 and 5 at their FULL sizes, checked through size-independent properties on the
device (the arrays never visit the host): the reference's known answer for config 2 (genuine
reference, row-major consolidate of the 2*10^8-entry generator => nnz 139,999,960, first entry (0, 1022331),
sum of values 200003308.953630), strict sortedness, idempotence; for config 5 the closed-form counts of the
pentadiagonal product, row-major sortedness, and exact linearity in a power-of-two scale."""
import numpy as np
import pytest

from _gpu import DevView

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


def test_config5_full_size(ctx):
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
    del i, k, v, i4, k4, v4
    for x in (A, B, w, C1, C4):
        x.free()


def test_config3_full_size(ctx):
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
    del i, k, v, it, kt, vt
    for x in (A, s, C1, Cs):
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
