#!/bin/bash
# round 2, call 2: the new parity tests (full-size oracle row samples, row panels, config 4 at scale 24)
set -u
out=gpurun_out/r02_c2
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run t_panels 300 python -m pytest tests/test_gpu_multiply.py -k "panels" -x -q -p no:cacheprovider
run t_nine 200 python -m pytest tests/test_gpu_consolidate.py -k "nine_bit" -x -q -p no:cacheprovider
run t_full 900 python -m pytest tests/test_gpu_full_size.py -x -q -s -p no:cacheprovider --durations=0
