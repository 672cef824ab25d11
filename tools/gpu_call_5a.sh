#!/bin/bash
# round 2, call 5a (1 GPU): radix passes with cp.async.bulk tile loads -- parity, then same-box A/B of the headline step
set -u
out=gpurun_out/r02_c5a
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run t_cons 600 python -m pytest tests/test_gpu_consolidate.py -x -q -p no:cacheprovider
run t_full 600 python -m pytest tests/test_gpu_full_size.py -x -q -p no:cacheprovider -k "config2 or config5"
SPB_BULK_LOAD=1 run bench_bulk1 300 python bench.py --no-e2e --no-cpu --no-also --steps 5 --warmup 3
SPB_BULK_LOAD=0 run bench_bulk0 300 python bench.py --no-e2e --no-cpu --no-also --steps 5 --warmup 3
SPB_BULK_LOAD=1 run probe1 100 python tools/radix9_probe.py 1e8 3
SPB_BULK_LOAD=0 run probe0 100 python tools/radix9_probe.py 1e8 3
