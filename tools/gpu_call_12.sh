#!/bin/bash
# round 2, call 12 (2 GPUs): where do 11 ms per step go at N=2 after the shard buffers were doubled?  (allocation trace)
set -u
out=gpurun_out/r02_c12
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
SPB_TRACE=1 run trace 300 $TR --master-port 29512 bench.py --gpus 2 --no-e2e --no-cpu --no-also --steps 4 --warmup 3
run plain 300 $TR --master-port 29513 bench.py --gpus 2 --no-e2e --no-cpu --no-also --steps 5 --warmup 3
