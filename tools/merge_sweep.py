"""Occupancy / variant sweep of the register-merge kernels on the config 3 and config 5 families:
python tools/merge_sweep.py [regrid|banded] [scale]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spsparse_b200 as sp


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "regrid"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    with sp.Context(0) as ctx:
        if what == "regrid":
            A = sp.gen_regrid(ctx, 0x5EED0003, int(3200 * scale), 3125, int(1000 * scale), 1000)
            s = sp.gen_vector(ctx, 0x5EED0013, int(1000 * scale) * 1000)
            Ar, Bt = sp.consolidate(ctx, A, sp.ROW_MAJOR), sp.consolidate(ctx, A, sp.COL_MAJOR)
            args = (1.0, None, Ar, 0, s, Bt, 1, None)
            fit = 4
        else:
            m = int(100_000_000 * scale)
            A, B, w = sp.gen_banded(ctx, 0x5EED0005, m, 0, m), sp.gen_banded(ctx, 0x5EED0015, m, 0, m), sp.gen_vector(ctx, 0x5EED0025, m)
            Ac, Bc = sp.consolidate(ctx, A, sp.ROW_MAJOR), sp.consolidate(ctx, B, sp.ROW_MAJOR)
            A.free(); B.free()
            args = (1.0, None, Ac, 0, w, Bc, 0, None)
            fit = 6

        def run():
            best = None
            for _ in range(3):
                C, st = sp.multiply_prepared(ctx, *args)
                C.free()
                t = (st.ms_symbolic, st.ms_numeric)
                best = t if best is None else (min(best[0], t[0]), min(best[1], t[1]))
            return best
        run()
        mode = sys.argv[3] if len(sys.argv) > 3 else "blocks"
        for nl in (fit, 8):
            for v in ((0, 3, 4, 5, 6, 7, 8, 10, 12) if mode == "blocks" else (0, 12, 25, 37, 50, 62, 75, 87, 100)):
                os.environ["SPB_MERGE_NL_COUNT"] = os.environ["SPB_MERGE_NL_NUMERIC"] = str(nl)
                key = "BLOCKS" if mode == "blocks" else "CARVEOUT"
                os.environ[f"SPB_MERGE_{key}_COUNT"] = os.environ[f"SPB_MERGE_{key}_NUMERIC"] = str(v)
                sym, num = run()
                print(f"{what} NL {nl} {mode} {v:3d}: symbolic {sym:7.3f} ms  numeric {num:7.3f} ms", flush=True)

main()
