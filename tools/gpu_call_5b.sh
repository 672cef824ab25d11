#!/bin/bash
# round 2, call 5b (2 GPUs): row-partitioned multiply with the own shard in place -- parity, then the N=2 line
set -u
out=gpurun_out/r02_c5b
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run check 400 $TR --master-port 29511 tests/multi/rowpart_check.py
run bench_n2 400 $TR --master-port 29512 bench.py --gpus 2 --no-e2e --no-cpu --steps 5 --warmup 3
