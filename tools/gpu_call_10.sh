#!/bin/bash
# round 2, call 10 (1 GPU): ncu --set full of the consolidate kernels as they are now (bulk-load passes, 128-bit histogram,
# in-row sort, reduce) on a 2.5e8-entry banded block, one launch of each
set -u
out=gpurun_out/r02_c10
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
cat > /tmp/cons_target.py <<'PY'
import sys
sys.path.insert(0, ".")
import spsparse_b200 as sp
with sp.Context(0) as ctx:
    A = sp.gen_banded(ctx, 0x5EED0005, 100_000_000, 0, 50_000_000)
    for _ in range(2):
        R, st = sp.consolidate(ctx, A, sp.ROW_MAJOR, stats=True)
        R.free()
    print(st.n_in, st.n_out, st.passes, st.digit_bits, st.ms_total, st.ms_pass, st.ms_reduce)
PY
run plain 120 python /tmp/cons_target.py
run ncu_cons 900 ncu --set full --import-source on --clock-control none -k regex:"k_radix_pass9|k_sort_hist_v4|k_segment_sort_walk|k_reduce_by_key" -s 7 -c 7 -o "$out/consolidate_kernels" python /tmp/cons_target.py
