#!/bin/bash
# round 2, call 26 (1 GPU): ncu --set full of every kernel of one consolidate (2.5e8-entry banded block), source-level
set -u
out=gpurun_out/r02_c26
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run list 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches.csv" python tools/env_ab_probe.py SPB_REDUCE_WARP 1 --rows 5e7 --iters 2 --no-config2
grep -c . "$out/launches.csv"
run ncu_all 900 ncu --set full --clock-control none --import-source on -k regex:"k_radix_pass9|k_segment_sort_walk|k_reduce_warp|k_sort_hist" -s 7 -c 7 -o "$out/consolidate" -f python tools/env_ab_probe.py SPB_REDUCE_WARP 1 --rows 5e7 --iters 2 --no-config2
ls -la "$out"
