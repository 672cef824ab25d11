#!/bin/bash
# round 2, call 20 (1 GPU): look-back rounds of 128 / 256 predecessors instead of 32 (reduce pass, scans, row heads) -- parity, A/B
set -u
out=gpurun_out/r02_c20
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
L=$PWD/spsparse_b200/lib
run t_cons 600 python -m pytest tests/test_gpu_consolidate.py tests/test_gpu_dense_ops.py -x -q -p no:cacheprovider
run base 120 python tools/radix9_probe.py 1e8 4
run base_c2 120 python tools/profile_target.py consolidate 1 4
for v in lb1_8 lb8_8 lb4_16; do
    SPB_LIB=$L/libspb_$v.so run $v 120 python tools/radix9_probe.py 1e8 4
    SPB_LIB=$L/libspb_$v.so run ${v}_c2 120 python tools/profile_target.py consolidate 1 4
done
