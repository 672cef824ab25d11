"""A/B of the separate in-row sort + reduce kernels against k_reduce_segsort (SPB_FUSED_REDUCE=1) on one banded block:
    python tools/fused_probe.py [rows=20000000] [iters=3]
Prints the consolidate phases of both variants (same process, same input) and checks that the outputs are identical."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import spsparse_b200 as sp  # noqa: E402


def main():
    m = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    with sp.Context(0) as ctx:
        A = sp.gen_banded(ctx, 0x5EED0005, m, 0, m)
        res = {}
        for fused in ("0", "1"):
            os.environ["SPB_FUSED_REDUCE"] = fused
            best = None
            for _ in range(iters):
                R, st = sp.consolidate(ctx, A, sp.ROW_MAJOR, stats=True)
                if best is None or st.ms_total < best.ms_total:
                    best = st
                if _ + 1 < iters:
                    R.free()
            idx, val = R.to_host()
            R.free()
            res[fused] = (idx, val)
            print(f"fused={fused}: n_in {best.n_in} n_out {best.n_out} passes {best.passes} total {best.ms_total:.3f} ms "
                  f"(sort incl. in-row sort {best.ms_sort:.3f}, pass {best.ms_pass:.3f}, reduce {best.ms_reduce:.3f})")
        same = all(np.array_equal(a, b) for a, b in zip(res["0"][0], res["1"][0])) and np.array_equal(res["0"][1], res["1"][1])
        print("outputs identical:", same)
        assert same


if __name__ == "__main__":
    main()
