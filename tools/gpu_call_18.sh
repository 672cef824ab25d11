#!/bin/bash
# round 2, call 18 (1 GPU box, mostly its CPU): smoke(), then the CPU baselines at BASELINE.md section 5's full sizes
set -u
out=gpurun_out/r02_c18
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run smoke 300 python __graft_entry__.py smoke
run cpu_full 1200 python bench.py --cpu-only --cpu-full
