#!/bin/bash
# round 2, call 32 (1 GPU): the full default bench line (all legs), the reference arm, the launch list of the bench command
set -u
out=gpurun_out/r02_c32
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run bench 1800 python bench.py
tail -c 300 "$out/bench.err"
run ref 900 python bench.py --impl reference --steps 2 --warmup 1
tail -c 300 "$out/ref.out"
run launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches.csv" python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-also
grep -c . "$out/launches.csv"
