#!/bin/bash
# round 2, call 27 (1 GPU): the whole GPU suite with the new reduce / walk / hash-walk paths, then a short bench line
set -u
out=gpurun_out/r02_c27
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run t_all 2400 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=8
tail -n 14 "$out/t_all.out"
run bench 900 python bench.py --steps 5 --warmup 3 --no-cpu --rmat-scale-named 0
tail -c 600 "$out/bench.err"
python - <<'P'
import json
for line in open("gpurun_out/r02_c27/bench.out"):
    if line.startswith("{"):
        d = json.loads(line)
        print("ms_per_step", d["ms_per_step"], "value", d["value"], "phases", d.get("phases_rank0", {}).get("timeline_ms"))
        print("roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "ms_per_launch")})
        a = d.get("also", {})
        for k, v in a.items():
            print(k, {kk: v[kk] for kk in ("ms_kernels", "ms_sort", "ms_reduce", "ms_symbolic", "ms_numeric", "model_frac") if kk in v})
        print("e2e", {k: d["e2e"][k] for k in ("ms_per_step", "value")})
P
