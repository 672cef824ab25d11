#!/bin/bash
# round 2, call 23 (1 GPU): ncu --set full of k_reduce_warp and k_reduce_by_key on a 2.5e8-entry banded block (source-level stalls)
set -u
out=gpurun_out/r02_c23
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run plain 300 python tools/env_ab_probe.py SPB_REDUCE_WARP 0 1 --rows 5e7 --iters 2 --no-config2
SPB_REDUCE_WARP=1 run ncu_warp 600 ncu --set full --clock-control none --import-source on -k regex:k_reduce_warp -s 1 -c 1 -o "$out/reduce_warp" -f python tools/env_ab_probe.py SPB_REDUCE_WARP 1 --rows 5e7 --iters 2 --no-config2
SPB_REDUCE_WARP=0 run ncu_block 600 ncu --set full --clock-control none --import-source on -k regex:k_reduce_by_key -s 1 -c 1 -o "$out/reduce_block" -f python tools/env_ab_probe.py SPB_REDUCE_WARP 0 --rows 5e7 --iters 2 --no-config2
ls -la "$out"
