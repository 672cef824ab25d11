"""R-MAT A*A (BASELINE config 4 family) timing probe: python tools/rmat_probe.py SCALE [variant ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    sc = int(sys.argv[1]) if len(sys.argv) > 1 else 18
    variants = [int(v) for v in sys.argv[2:]] or [0, 1]
    import spsparse_b200 as sp
    for var in variants:
        os.environ["SPB_HASH_VARIANT"] = str(var)
        with sp.Context(0) as ctx:
            A = sp.gen_rmat(ctx, 0x5EED0004, sc, 4 << sc)
            Ac = sp.consolidate(ctx, A, (0, 1))
            best = None
            for it in range(3):
                if it == 2 and os.environ.get("SPB_TRACE"):
                    sys.stderr.write("---- last iteration\n")
                Cm, st = sp.multiply_prepared(ctx, 1.0, None, Ac, 0, None, Ac, 0, None)
                d = st.asdict()
                Cm.free()
                if it and (best is None or d["ms_symbolic"] + d["ms_numeric"] < best["ms_symbolic"] + best["ms_numeric"]):
                    best = d
            print("variant", var, "scale", sc, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in best.items()}, flush=True)
            A.free(); Ac.free()


main()
