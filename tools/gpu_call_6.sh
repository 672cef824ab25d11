#!/bin/bash
# round 2, call 6 (1 GPU): the whole GPU suite, then the default bench line (all legs)
set -u
out=gpurun_out/r02_c6
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run t_all 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=8
run bench_n1 1500 python bench.py
run bench_ref 600 python bench.py --impl reference
