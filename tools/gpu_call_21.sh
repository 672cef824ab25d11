#!/bin/bash
# round 2, call 21 (1 GPU): hash bin -- bitmap walk over the non-zero words only (summary bits): parity, A/B at scale 20 and at 2^24 rows
set -u
out=gpurun_out/r02_c21
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run t_mult 900 python -m pytest tests/test_gpu_multiply.py -x -q -p no:cacheprovider
tail -n 3 "$out/t_mult.out"
SPB_HASH_SPARSE_WALK=0 run rmat20_walk0 200 python tools/rmat_probe.py 20 2
SPB_HASH_SPARSE_WALK=1 run rmat20_walk1 200 python tools/rmat_probe.py 20 2
cat "$out/rmat20_walk0.out" "$out/rmat20_walk1.out"
run named 600 python tools/rmat_named_probe.py 24 10 0 1
cat "$out/named.out"
run t_full4 900 python -m pytest tests/test_gpu_full_size.py -x -q -p no:cacheprovider -k "config4_row_sample"
tail -n 3 "$out/t_full4.out"
