#!/bin/bash
# round 2, call 30 (2 GPUs): the row-partitioned path with the new reduce / in-row sort kernels -- parity under torchrun, bench N=2
set -u
out=gpurun_out/r02_c30
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run t_two 900 python -m pytest tests/test_gpu_multiply.py -x -q -p no:cacheprovider -k "two_gpus or one_rank"
tail -n 3 "$out/t_two.out"
run bench2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3
tail -c 400 "$out/bench2.err"
python - <<'P'
import json
for line in open("gpurun_out/r02_c30/bench2.out"):
    if line.startswith("{"):
        d = json.loads(line)
        print("N=2 ms_per_step", d["ms_per_step"], "value", d["value"], "timeline", d.get("phases_rank0", {}).get("timeline_ms"))
        print("fingerprint", d["config"]["result_fingerprint"], "e2e", d["e2e"]["ms_per_step"], "full_replicate", d.get("also", {}).get("full_replicate", {}).get("ms_per_step"))
P
python - <<'P'
import json
for line in open("gpurun_out/r02_c30/bench2.out"):
    if line.startswith("{"):
        d = json.loads(line)
        print("per_rank", json.dumps(d.get("per_rank")))
P
