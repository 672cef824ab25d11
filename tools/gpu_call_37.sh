#!/bin/bash
# round 2, call 37 (1 GPU): value-free symbolic pass + tombstones written in place by the numeric pass -- parity, A/B against the build before the change
set -u
out=gpurun_out/r02_c37
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
L=$PWD/spsparse_b200/lib
run t_mult 900 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_dropin.py -x -q -p no:cacheprovider
tail -n 3 "$out/t_mult.out"
for rep in 1 2; do
    SPB_LIB=$L/libspb_prev0.so run prev0_b$rep 300 python tools/profile_target.py banded 1 4
    run new_b$rep 300 python tools/profile_target.py banded 1 4
    SPB_LIB=$L/libspb_prev0.so run prev0_r$rep 300 python tools/profile_target.py regrid 1 4
    run new_r$rep 300 python tools/profile_target.py regrid 1 4
    cat "$out/prev0_b$rep.out" "$out/new_b$rep.out" "$out/prev0_r$rep.out" "$out/new_r$rep.out"
done
run t_full 900 python -m pytest tests/test_gpu_full_size.py -x -q -p no:cacheprovider -k "config5 or config3 or config4_row_sample"
tail -n 3 "$out/t_full.out"
