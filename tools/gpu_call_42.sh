#!/bin/bash
# round 2, call 42 (2 GPUs): side stream of the row-partitioned multiply at default priority (SPB_ROWPART_SIDE_PRIO=0) against high priority
set -u
out=gpurun_out/r02_c42
mkdir -p "$out"
for v in 0 1; do
  SPB_ROWPART_SIDE_PRIO=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$v bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-also > "$out/b$v.out" 2> "$out/b$v.err"
  python - <<P
import json
for line in open("$out/b$v.out"):
    if line.startswith("{"):
        d = json.loads(line)
        print("prio=$v ms_per_step", d["ms_per_step"], json.dumps(d["per_rank"]["phases_ms"]))
P
done
