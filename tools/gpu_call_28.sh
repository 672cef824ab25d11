#!/bin/bash
# round 2, call 28 (1 GPU): in-row walk with row-start bit words, a warp per 256 consecutive entries -- parity, A/B against the previous build
set -u
out=gpurun_out/r02_c28
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
L=$PWD/spsparse_b200/lib
run t_cons 900 python -m pytest tests/test_gpu_consolidate.py -x -q -p no:cacheprovider
tail -n 3 "$out/t_cons.out"
SPB_LIB=$L/libspb_prev.so run prev 300 python tools/env_ab_probe.py SPB_REDUCE_WARP 1 --no-config2
run new 300 python tools/env_ab_probe.py SPB_REDUCE_WARP 1 --no-config2
SPB_LIB=$L/libspb_prev.so run prev2 300 python tools/env_ab_probe.py SPB_REDUCE_WARP 1 --no-config2
run new2 300 python tools/env_ab_probe.py SPB_REDUCE_WARP 1 --no-config2
cat "$out/prev.out" "$out/new.out" "$out/prev2.out" "$out/new2.out"
run t_full 900 python -m pytest tests/test_gpu_full_size.py -x -q -p no:cacheprovider -k "config2 or config5"
tail -n 3 "$out/t_full.out"
SPB_LIB=$L/libspb_prev.so run prev_c2 300 python tools/profile_target.py consolidate 1 4
run new_c2 300 python tools/profile_target.py consolidate 1 4
cat "$out/prev_c2.out" "$out/new_c2.out"
