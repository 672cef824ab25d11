#!/bin/bash
# round 2, call 33 (1 GPU): in-row sort on shuffles for rows of at most 5 entries (k_segment_sort_shfl) -- parity, same-process A/B
set -u
out=gpurun_out/r02_c33
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run t_cons 900 python -m pytest tests/test_gpu_consolidate.py -x -q -p no:cacheprovider
tail -n 3 "$out/t_cons.out"
run ab 600 python tools/env_ab_probe.py SPB_SEGMENT_WALK 1 2 1 2 --no-config2 --iters 3
cat "$out/ab.out"
run t_full 900 python -m pytest tests/test_gpu_full_size.py -x -q -p no:cacheprovider -k "config5 or config3 or config2"
tail -n 3 "$out/t_full.out"
