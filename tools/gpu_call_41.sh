#!/bin/bash
# round 2, call 41 (1 GPU): last check -- smoke(), the consolidate tests (incl. KEEP_ALL through the row-pass organisation), the multiply tests
set -u
out=gpurun_out/r02_c41
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
tail -n 1 "$out/smoke.out"
run t_cons 600 python -m pytest tests/test_gpu_consolidate.py tests/test_gpu_multiply.py tests/test_gpu_dropin.py tests/test_gpu_dense_ops.py -x -q -p no:cacheprovider
tail -n 3 "$out/t_cons.out"
