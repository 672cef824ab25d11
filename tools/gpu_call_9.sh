#!/bin/bash
# round 2, call 9 (1 GPU): parity of the multiply paths after the last hash-bin / pool changes, config 4 timings
set -u
out=gpurun_out/r02_c9
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run t_mult 600 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_dropin.py tests/test_gpu_dense_ops.py -x -q -p no:cacheprovider
run t_full 900 python -m pytest tests/test_gpu_full_size.py -x -q -p no:cacheprovider
run bench 900 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3
