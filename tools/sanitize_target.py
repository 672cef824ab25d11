"""Small run through every kernel family for compute-sanitizer:  compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    import spsparse_b200 as sp
    rng = np.random.default_rng(5)
    for walk in ("1", "0"):
        os.environ["SPB_SEGMENT_SORT"], os.environ["SPB_SEGMENT_WALK"] = "1", walk
        with sp.Context(0) as ctx:
            n = 20000
            i, k = rng.integers(0, 3000, n), rng.integers(0, 1 << 20, n)
            i[:700] = 17  # a row longer than the in-row limit: full-key fallback
            A = sp.CooArray.from_host(ctx, (3000, 1 << 20), [i, k], rng.standard_normal(n))
            for so in ((0, 1), (1, 0)):
                R = sp.consolidate(ctx, A, so)
                R.free()
            A.free()
    # no long rows: the reduce pass keeps the result it formed from the in-row sort's head counts (no look-back); "2": the
    # warp-per-tile kernel with its look-back, also after a full-key sort
    for rw, seg in (("1", "1"), ("2", "1"), ("2", "0"), ("0", "1")):
        os.environ["SPB_REDUCE_WARP"], os.environ["SPB_SEGMENT_SORT"] = rw, seg
        with sp.Context(0) as ctx:
            n = 30011
            i, k = rng.integers(0, 9000, n), rng.integers(0, 40, n)      # short rows, many duplicates
            A = sp.CooArray.from_host(ctx, (9000, 1 << 20), [i, k], rng.standard_normal(n))
            for so, pol in (((0, 1), 1), ((1, 0), 2), ((0, 1), 0)):
                R = sp.consolidate(ctx, A, so, pol)
                R.free()
            A.free()
    os.environ.pop("SPB_REDUCE_WARP")
    os.environ.pop("SPB_SEGMENT_SORT"); os.environ.pop("SPB_SEGMENT_WALK")
    for env in ({}, {"SPB_MERGE_MAX_PRODUCTS": "0", "SPB_HASH_MIN_PRODUCTS": "0", "SPB_HASH_WIN_COLS": "64", "SPB_HASH_ITEM_CAP": "5"},
                {"SPB_MERGE_MAX_PRODUCTS": "0", "SPB_HASH_MIN_PRODUCTS": "0"},                                  # one window: sparse bitmap walk
                {"SPB_MERGE_MAX_PRODUCTS": "0", "SPB_HASH_MIN_PRODUCTS": "0", "SPB_HASH_SPARSE_WALK": "0"},      # ... and the group walk
                {"SPB_MERGE_MAX_PRODUCTS": "0", "SPB_HASH_MIN_PRODUCTS": "0", "SPB_HASH_WIN_COLS": "64", "SPB_HASH_SMALL": "1"},
                {"SPB_HASH_MIN_PRODUCTS": "off", "SPB_MERGE_MAX_PRODUCTS": "0", "SPB_ESC_CHUNK": "500"}):
        os.environ.update(env)
        with sp.Context(0) as ctx:
            A = sp.gen_rmat(ctx, 0x5EED0004, 9, 4 << 9)
            C, st = sp.multiply(ctx, 1.0, None, A, ".", None, A, ".", None, stats=True)
            d = sp.to_dense(ctx, C)
            S = sp.to_sparse(ctx, d)
            T = sp.transpose(ctx, S, (1, 0))
            for x in (A, C, S, T):
                x.free()
        for kx in env:
            os.environ.pop(kx)
    print("sanitize target done", st.asdict()["products"])


main()
