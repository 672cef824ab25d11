"""Small fixed workloads for ncu: `python tools/profile_target.py consolidate|regrid|banded [scale]`.
Runs the same library calls bench.py times, a few iterations, nothing else."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spsparse_b200 as sp  # noqa: E402


def main():
    what = sys.argv[1]
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    with sp.Context(0) as ctx:
        if what == "consolidate":
            n = int(200_000_000 * scale)
            A = sp.gen_dup_coo(ctx, 0x5EED0002, 0, n, int(n * 0.7), 24, 0)
            for _ in range(iters):
                R, st = sp.consolidate(ctx, A, sp.ROW_MAJOR, stats=True)
                R.free()
            print("consolidate", n, "->", st.n_out, f"{st.ms_total:.3f} ms (sort {st.ms_sort:.3f}, pass {st.ms_pass:.3f}, reduce {st.ms_reduce:.3f})")
        elif what == "regrid":
            ny, nx, gy, gx = int(3200 * scale), 3125, int(1000 * scale), 1000
            A = sp.gen_regrid(ctx, 0x5EED0003, ny, nx, gy, gx)
            s = sp.gen_vector(ctx, 0x5EED0013, gy * gx)
            Ar, Bt = sp.consolidate(ctx, A, sp.ROW_MAJOR), sp.consolidate(ctx, A, sp.COL_MAJOR)
            for _ in range(iters):
                C, st = sp.multiply_prepared(ctx, 1.0, None, Ar, 0, s, Bt, 1, None)
                C.free()
            print("regrid F", st.products, "nnzC", st.nnz_c, f"prep {st.ms_prepare:.3f} sym {st.ms_symbolic:.3f} num {st.ms_numeric:.3f} ms")
        elif what == "banded":
            m = int(100_000_000 * scale)
            A, B, w = sp.gen_banded(ctx, 0x5EED0005, m, 0, m), sp.gen_banded(ctx, 0x5EED0015, m, 0, m), sp.gen_vector(ctx, 0x5EED0025, m)
            Ac, Bc = sp.consolidate(ctx, A, sp.ROW_MAJOR), sp.consolidate(ctx, B, sp.ROW_MAJOR)
            A.free(); B.free()
            for _ in range(iters):
                C, st = sp.multiply_prepared(ctx, 1.0, None, Ac, 0, w, Bc, 0, None)
                C.free()
            print("banded F", st.products, "nnzC", st.nnz_c, f"prep {st.ms_prepare:.3f} sym {st.ms_symbolic:.3f} num {st.ms_numeric:.3f} ms")
        elif what == "rmat":
            sc = int(sys.argv[2]) if len(sys.argv) > 2 else 20
            A = sp.gen_rmat(ctx, 0x5EED0004, sc, 4 << sc)
            for _ in range(iters):
                C, st = sp.multiply(ctx, 1.0, None, A, ".", None, A, ".", None, stats=True)
                C.free()
            print("rmat", sc, st.asdict())


if __name__ == "__main__":
    main()
