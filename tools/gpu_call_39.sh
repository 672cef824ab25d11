#!/bin/bash
# round 2, call 39 (1 GPU): whole GPU suite, then the full default bench line, reference arm, launch list -- final state
set -u
out=gpurun_out/r02_c39
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
L=$PWD/spsparse_b200/lib
SPB_LIB=$L/libspb_prev0.so run prev0_b 300 python tools/profile_target.py banded 1 4
run new_b 300 python tools/profile_target.py banded 1 4
cat "$out/prev0_b.out" "$out/new_b.out"
run t_all 2400 python -m pytest tests -m gpu -x -q -p no:cacheprovider
tail -n 3 "$out/t_all.out"
run bench 1800 python bench.py
tail -c 300 "$out/bench.err"
run ref 900 python bench.py --impl reference --steps 3 --warmup 3
run launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches.csv" python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-also
grep -c . "$out/launches.csv"
python - <<'P'
import json
for line in open("gpurun_out/r02_c39/bench.out"):
    if line.startswith("{"):
        d = json.loads(line)
        print("ms_per_step", d["ms_per_step"], "value", d["value"], "timeline", d["phases_rank0"]["timeline_ms"])
        for k, v in d["also"].items():
            print(k, {kk: v[kk] for kk in ("ms_kernels", "ms_symbolic", "ms_numeric", "model_frac", "ms_sweep_wall") if kk in v})
        print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["cabi_pageable"]["ms_per_step"], d["e2e"]["cpp_api"]["ms_per_step"])
P
