"""A/B of 8-bit against 9-bit radix digits (SPB_RADIX9=1, k_radix_pass9) on one banded block (27-bit row part: 4 -> 3 passes):
    python tools/radix9_probe.py [rows=20000000] [iters=3]
Same process, same input; prints the consolidate phases of both and checks that the outputs are identical."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import spsparse_b200 as sp  # noqa: E402


def main():
    m = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    with sp.Context(0) as ctx:
        # the SHAPE decides the key width: a block of m rows of the 10^8-row matrix keeps the 27-bit row part
        A = sp.gen_banded(ctx, 0x5EED0005, 100_000_000, 0, m)
        res = {}
        for nine in ("0", "1"):
            os.environ["SPB_RADIX9"] = nine
            best = None
            for it in range(iters):
                R, st = sp.consolidate(ctx, A, sp.ROW_MAJOR, stats=True)
                if best is None or st.ms_total < best.ms_total:
                    best = st
                if it + 1 < iters:
                    R.free()
            idx, val = R.to_host()
            R.free()
            res[nine] = (idx, val)
            print(f"radix9={nine}: n_in {best.n_in} n_out {best.n_out} passes {best.passes} total {best.ms_total:.3f} ms "
                  f"(sort incl. in-row sort {best.ms_sort:.3f}, pass {best.ms_pass:.3f}, reduce {best.ms_reduce:.3f})", flush=True)
        same = all(np.array_equal(a, b) for a, b in zip(res["0"][0], res["1"][0])) and np.array_equal(res["0"][1], res["1"][1])
        print("outputs identical:", same)
        assert same


if __name__ == "__main__":
    main()
