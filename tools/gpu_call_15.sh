#!/bin/bash
# round 2, call 15 (1 GPU): one-pass merge for short-row matrices -- parity (all multiply tests run through it where it applies),
# same-process A/B, headline step
set -u
out=gpurun_out/r02_c15
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run t_mult 600 python -m pytest tests/test_gpu_multiply.py tests/test_gpu_dropin.py tests/test_gpu_dense_ops.py -x -q -p no:cacheprovider
run t_full 900 python -m pytest tests/test_gpu_full_size.py -x -q -p no:cacheprovider -k "config5 or config3 or row_sample"
run one_banded 200 python tools/profile_target.py banded 1 4
SPB_MERGE_ONEPASS=0 run two_banded 200 python tools/profile_target.py banded 1 4
run regrid 200 python tools/profile_target.py regrid 1 3
run rmat20 200 python tools/profile_target.py rmat 20 3
run bench 600 python bench.py --no-e2e --no-cpu --no-also --steps 5 --warmup 3
