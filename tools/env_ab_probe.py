"""Same-process A/B of one environment knob on the two consolidate workloads (banded block of BASELINE config 5, config 2):
    python tools/env_ab_probe.py KNOB VALUE [VALUE ...] [--rows R] [--iters N] [--no-config2]
Prints the phases of consolidate for every value and checks that all values give identical outputs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import spsparse_b200 as sp  # noqa: E402


def main():
    args = sys.argv[1:]
    rows, iters, c2 = 100_000_000, 4, True
    if "--rows" in args:
        i = args.index("--rows"); rows = int(float(args[i + 1])); del args[i:i + 2]
    if "--iters" in args:
        i = args.index("--iters"); iters = int(args[i + 1]); del args[i:i + 2]
    if "--no-config2" in args:
        args.remove("--no-config2"); c2 = False
    knob, values = args[0], args[1:]
    with sp.Context(0) as ctx:
        work = [("banded", lambda: sp.gen_banded(ctx, 0x5EED0005, 100_000_000, 0, rows))]
        if c2:
            work.append(("config2", lambda: sp.gen_dup_coo(ctx, 0x5EED0002, 0, 200_000_000, 140_000_000, 24, 0)))
        for name, make in work:
            A = make()
            ref = None
            for v in values:
                if v == "unset":
                    os.environ.pop(knob, None)
                else:
                    os.environ[knob] = v
                best = None
                for it in range(iters):
                    R, st = sp.consolidate(ctx, A, sp.ROW_MAJOR, stats=True)
                    if best is None or st.ms_total < best.ms_total:
                        best = st
                    if it + 1 < iters:
                        R.free()
                idx, val = R.to_host()
                R.free()
                print(f"{name} {knob}={v}: n_in {best.n_in} n_out {best.n_out} passes {best.passes} total {best.ms_total:.3f} ms "
                      f"(sort incl. in-row sort {best.ms_sort:.3f}, pass {best.ms_pass:.3f}, reduce {best.ms_reduce:.3f})", flush=True)
                if ref is None:
                    ref = (idx, val)
                else:
                    same = all(np.array_equal(a, b) for a, b in zip(ref[0], idx)) and np.array_equal(ref[1], val)
                    print("  identical to the first value's output:", same, flush=True)
                    assert same
                    del idx, val
            del ref
            A.free()
            ctx.trim()


if __name__ == "__main__":
    main()
