#!/bin/bash
# round 2, call 34 (1 GPU): value-free symbolic pass of the register-merge bin (tombstones for cancelled outputs) -- parity, same-process A/B
set -u
out=gpurun_out/r02_c34
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run t_mult 900 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_dropin.py -x -q -p no:cacheprovider
tail -n 3 "$out/t_mult.out"
for v in 1 0 1 0; do
    SPB_MERGE_EXACT_COUNT=$v run banded_$v 300 python tools/profile_target.py banded 1 4
    SPB_MERGE_EXACT_COUNT=$v run regrid_$v 300 python tools/profile_target.py regrid 1 4
    echo "exact=$v"; cat "$out/banded_$v.out" "$out/regrid_$v.out"
done
run t_full 900 python -m pytest tests/test_gpu_full_size.py -x -q -p no:cacheprovider -k "config5 or config3"
tail -n 3 "$out/t_full.out"
