#!/bin/bash
# round 2, call 7 (1 GPU): hash-bin restructure -- parity, then config 4 timings (scale 20 and as named) and the headline step
set -u
out=gpurun_out/r02_c7
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run t_mult 900 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_dropin.py -x -q -p no:cacheprovider
run t_full4 900 python -m pytest tests/test_gpu_full_size.py -x -q -p no:cacheprovider -k "config4 or config5 or config3"
run bench 900 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3
