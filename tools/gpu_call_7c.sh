#!/bin/bash
# round 2, call 7c (1 GPU): merge kernels after the fixes (per-warp striped counters, 28 KB block path), same-box A/B
set -u
out=gpurun_out/r02_c7c
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
OLD=$PWD/spsparse_b200/lib/libspsparse_b200_cb5300b.so
for w in banded regrid; do
    run new_$w 200 python tools/profile_target.py $w 1 4
    SPB_LIB=$OLD run old_$w 200 python tools/profile_target.py $w 1 4
    SPB_MERGE_DEBUG=3 run newnoblock_$w 200 python tools/profile_target.py $w 1 4
done
run t_mult 600 python -m pytest tests/test_gpu_multiply.py -x -q -p no:cacheprovider
run bench 600 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3
