#!/bin/bash
# round 2, call 19 (1 GPU): radix-pass tile size / blocks per SM -- same-box A/B (16 entries per thread x 3 blocks is the default)
set -u
out=gpurun_out/r02_c19
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
L=$PWD/spsparse_b200/lib
run base 120 python tools/radix9_probe.py 1e8 4
run base_c2 120 python tools/profile_target.py consolidate 1 4
for v in ipt24_2 ipt20_2; do
    SPB_LIB=$L/libspb_$v.so run $v 120 python tools/radix9_probe.py 1e8 4
    SPB_LIB=$L/libspb_$v.so run ${v}_c2 120 python tools/profile_target.py consolidate 1 4
done
