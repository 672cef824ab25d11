"""Config 4 as named (R-MAT 2^SCALE rows, A*A in row panels): per-kernel times of the first PANELS panels, for every value of
SPB_HASH_SPARSE_WALK given.  python tools/rmat_named_probe.py SCALE PANELS [walk ...]   (walk: 0 = group walk, 1 = sparse walk)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    sc = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    panels = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    walks = sys.argv[3:] or ["0", "1"]
    import spsparse_b200 as sp
    keys = ("ms_merge_count", "ms_hash_count", "ms_esc", "ms_merge_numeric", "ms_hash_emit", "ms_hash_splits", "ms_hash_numeric", "ms_prepare")
    with sp.Context(0) as ctx:
        A = sp.gen_rmat(ctx, 0x5EED0004, sc, 4 << sc)
        plan = sp.MultiplyPlan(ctx, 1.0, None, A, ".", None, A, ".", None, max_products_per_panel=1 << 30)
        A.free()
        n = min(panels, plan.n_panels) if panels > 0 else plan.n_panels
        # panels spread over the sweep (the first ones hold the hub rows, the last ones the short rows)
        pick = sorted(set(int(i * plan.n_panels / n) for i in range(n)))
        for rep in range(2):
            for walk in walks:
                os.environ["SPB_HASH_SPARSE_WALK"] = walk
                ctx.sync()
                t0 = time.perf_counter()
                tot = {k: 0.0 for k in keys}
                prod = nnz = 0
                for p in pick:
                    Cp, st = plan.panel(p, stats=True)
                    for k in keys:
                        tot[k] += getattr(st, k)
                    prod += st.products; nnz += st.nnz_c
                    Cp.free()
                ctx.sync()
                wall = time.perf_counter() - t0
                print(f"rep {rep} walk {walk} scale {sc} panels {len(pick)}/{plan.n_panels} wall {wall * 1e3:.1f} ms products {prod} nnz_c {nnz}",
                      {k: round(v, 2) for k, v in tot.items()}, flush=True)
        plan.free()


main()
