#!/bin/bash
# round 2, call 7b (1 GPU): why did the merge kernels' phases get slower between cb5300b and dc00636?  Same box, both libraries.
set -u
out=gpurun_out/r02_c7b
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
OLD=$PWD/spsparse_b200/lib/libspsparse_b200_cb5300b.so
for w in banded regrid; do
    run new_$w 200 python tools/profile_target.py $w 1 4
    SPB_LIB=$OLD run old_$w 200 python tools/profile_target.py $w 1 4
    SPB_TRACE=1 run newtrace_$w 200 python tools/profile_target.py $w 1 3
    SPB_MERGE_DEBUG=3 run newnoblock_$w 200 python tools/profile_target.py $w 1 4
done
run rmat20 200 python tools/profile_target.py rmat 20 3
run t_cons 600 python -m pytest tests/test_gpu_consolidate.py -x -q -p no:cacheprovider
run t_mult 600 python -m pytest tests/test_gpu_multiply.py -x -q -p no:cacheprovider
run bench 600 python bench.py --no-e2e --no-cpu --no-also --steps 5 --warmup 3
