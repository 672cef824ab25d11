#!/bin/bash
# round 2, call 11 (1 GPU): reduce pass with 128-thread blocks (fewer warps per barrier, more blocks per SM) -- same-box A/B
set -u
out=gpurun_out/r02_c11
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
L=$PWD/spsparse_b200/lib
run base 120 python tools/radix9_probe.py 1e8 4
for v in rk128_10 rk128_8 rk256_6; do
    SPB_LIB=$L/libspb_$v.so run $v 120 python tools/radix9_probe.py 1e8 4
done
run cons2 120 python tools/profile_target.py consolidate 1 4
SPB_LIB=$L/libspb_rk128_10.so run cons2_rk128_10 120 python tools/profile_target.py consolidate 1 4
