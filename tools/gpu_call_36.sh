#!/bin/bash
# round 2, call 36 (1 GPU): numeric merge pass before / after the tombstone check (three builds), same box
set -u
out=gpurun_out/r02_c36
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
L=$PWD/spsparse_b200/lib
for rep in 1 2; do
    SPB_LIB=$L/libspb_prev0.so run prev0_b$rep 300 python tools/profile_target.py banded 1 4
    SPB_LIB=$L/libspb_prev.so run prev_b$rep 300 python tools/profile_target.py banded 1 4
    run new_b$rep 300 python tools/profile_target.py banded 1 4
    cat "$out/prev0_b$rep.out" "$out/prev_b$rep.out" "$out/new_b$rep.out"
done
