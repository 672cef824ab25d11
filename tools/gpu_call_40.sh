#!/bin/bash
# round 2, call 40 (8 GPUs): bench N=8 with the final kernels (headline + full-replicate + end-to-end), rowpart parity under torchrun on 8 ranks
set -u
out=gpurun_out/r02_c40
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run bench8 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3
tail -c 300 "$out/bench8.err"
python - <<'P'
import json
for line in open("gpurun_out/r02_c40/bench8.out"):
    if line.startswith("{"):
        d = json.loads(line)
        print("N=8 ms_per_step", d["ms_per_step"], "value", d["value"], "timeline", d.get("phases_rank0", {}).get("timeline_ms"))
        print("fingerprint", d["config"]["result_fingerprint"], "e2e", d["e2e"]["ms_per_step"], "full_replicate", d.get("also", {}).get("full_replicate", {}).get("ms_per_step"))
P


python - <<'P'
import json
for line in open("gpurun_out/r02_c40/bench8.out"):
    if line.startswith("{"):
        d = json.loads(line)
        print("per_rank", json.dumps(d.get("per_rank")))
P
