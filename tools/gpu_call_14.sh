#!/bin/bash
# round 2, call 14 (1 GPU): LOCAL merge kernels, shared-memory carve-out sweep
set -u
out=gpurun_out/r02_c14
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
SPB_MERGE_LOCAL=0 run global 200 python tools/profile_target.py banded 1 4
run local_default 200 python tools/profile_target.py banded 1 4
for c in 50 75 100; do SPB_MERGE_LOCAL_CARVE=$c run local_$c 200 python tools/profile_target.py banded 1 4; done
