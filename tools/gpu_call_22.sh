#!/bin/bash
# round 2, call 22 (1 GPU): duplicate-reduce with a warp per tile (k_reduce_warp) -- parity, same-process A/B
set -u
out=gpurun_out/r02_c22
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run t_cons 900 python -m pytest tests/test_gpu_consolidate.py tests/test_gpu_dense_ops.py -x -q -p no:cacheprovider
tail -n 3 "$out/t_cons.out"
run ab 600 python tools/env_ab_probe.py SPB_REDUCE_WARP 0 1
cat "$out/ab.out"
