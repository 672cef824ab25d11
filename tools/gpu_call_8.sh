#!/bin/bash
# round 2, call 8 (8 GPUs): the row-partitioned multiply at 8 ranks -- parity, the N=8 line (with e2e and the fetch-all line),
# and the round-1 path beside it
set -u
out=gpurun_out/r02_c8
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run check 300 $TR --master-port 29511 tests/multi/rowpart_check.py
run bench_n8 600 $TR --master-port 29512 bench.py --gpus 8 --no-cpu --steps 5 --warmup 3
SPB_LEGACY_DIST=1 run bench_n8_legacy 300 $TR --master-port 29513 bench.py --gpus 8 --no-e2e --no-cpu --no-also --steps 5 --warmup 3
