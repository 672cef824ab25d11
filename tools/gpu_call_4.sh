#!/bin/bash
# round 2, call 4 (2 GPUs): the row-partitioned multiply behind the C ABI -- parity under torchrun, then N=2 bench lines
# (new path, legacy path), and the N=1 line through the same entry point
set -u
out=gpurun_out/r02_c4
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run check 400 $TR --master-port 29511 tests/multi/rowpart_check.py
run t_one 200 python -m pytest tests/test_gpu_multiply.py -k "one_rank or two_gpus" -x -q -p no:cacheprovider
run bench_n2 400 $TR --master-port 29512 bench.py --gpus 2 --no-e2e --no-cpu --steps 5 --warmup 3
SPB_LEGACY_DIST=1 run bench_n2_legacy 400 $TR --master-port 29513 bench.py --gpus 2 --no-e2e --no-cpu --no-also --steps 5 --warmup 3
run bench_n1 400 python bench.py --no-e2e --no-cpu --no-also --steps 5 --warmup 3
