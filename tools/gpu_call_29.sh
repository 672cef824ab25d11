#!/bin/bash
# round 2, call 29 (1 GPU): in-row walk loops unrolled (4 / 8 / 16) against the rolled loops -- parity, same-box A/B
set -u
out=gpurun_out/r02_c29
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
L=$PWD/spsparse_b200/lib
run t_cons 900 python -m pytest tests/test_gpu_consolidate.py -x -q -p no:cacheprovider
tail -n 3 "$out/t_cons.out"
for rep in 1 2; do
for v in walk3 unroll4 unroll16; do
    SPB_LIB=$L/libspb_$v.so run ${v}_$rep 300 python tools/env_ab_probe.py SPB_REDUCE_WARP 1 --no-config2 --iters 3
    echo $v; cat "$out/${v}_$rep.out"
done
run unroll8_$rep 300 python tools/env_ab_probe.py SPB_REDUCE_WARP 1 --no-config2 --iters 3
echo unroll8; cat "$out/unroll8_$rep.out"
done
