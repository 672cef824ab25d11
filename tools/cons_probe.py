"""consolidate timing probe (configs 2 and 5 inputs): python tools/cons_probe.py  [SPB_LIB=other.so for A/B runs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spsparse_b200 as sp
with sp.Context(0) as ctx:
    for what in ("cfg2", "banded"):
        if what == "cfg2":
            A = sp.gen_dup_coo(ctx, 0x5EED0002, 0, 200_000_000, 140_000_000, 24, 0)
        else:
            A = sp.gen_banded(ctx, 0x5EED0005, 100_000_000, 0, 100_000_000)
        best = None
        for _ in range(5):
            R, st = sp.consolidate(ctx, A, sp.ROW_MAJOR, stats=True)
            R.free()
            t = (st.ms_total, st.ms_sort, st.ms_pass, st.ms_reduce)
            best = t if best is None or t[0] < best[0] else best
        print(what, "total %.3f sort %.3f pass %.3f reduce %.3f" % best, flush=True)
        A.free()
