#!/bin/bash
# round 2, call 35 (1 GPU): value-free symbolic pass, numeric with the tombstone check after the merge -- parity (incl. the new cancellation test), A/B against the previous build
set -u
out=gpurun_out/r02_c35
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
L=$PWD/spsparse_b200/lib
run t_mult 900 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_dropin.py -x -q -p no:cacheprovider
tail -n 3 "$out/t_mult.out"
for rep in 1 2; do
    SPB_LIB=$L/libspb_prev.so run prev_b$rep 300 python tools/profile_target.py banded 1 4
    run new_b$rep 300 python tools/profile_target.py banded 1 4
    SPB_LIB=$L/libspb_prev.so run prev_r$rep 300 python tools/profile_target.py regrid 1 4
    run new_r$rep 300 python tools/profile_target.py regrid 1 4
    cat "$out/prev_b$rep.out" "$out/new_b$rep.out" "$out/prev_r$rep.out" "$out/new_r$rep.out"
done
