#!/bin/bash
# round 2, call 3: tests of the 2^31 cap / ordered long runs / scale check / trim, and the first timings of config 4 at scale 24
set -u
out=gpurun_out/r02_c3
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run t_cons 600 python -m pytest tests/test_gpu_consolidate.py -x -q -p no:cacheprovider --durations=5
run t_mult 600 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_dense_ops.py tests/test_gpu_dropin.py -x -q -p no:cacheprovider --durations=5
run bench_also 900 python bench.py --rows 2000000 --no-e2e --no-cpu --steps 2 --warmup 3
