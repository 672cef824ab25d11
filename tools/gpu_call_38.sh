#!/bin/bash
# round 2, call 38 (1 GPU): kernel durations of the merge passes, build before the value-free symbolic pass vs now (ncu launch list)
set -u
out=gpurun_out/r02_c38
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
L=$PWD/spsparse_b200/lib
SPB_LIB=$L/libspb_prev0.so run prev0 300 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_merge --csv --log-file "$out/prev0.csv" python tools/profile_target.py banded 1 2
run new 300 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_merge --csv --log-file "$out/new.csv" python tools/profile_target.py banded 1 2
python - <<'P'
import csv
for f in ("prev0","new"):
    rows=[r for r in csv.reader(open(f"gpurun_out/r02_c38/{f}.csv")) if len(r)>8]
    h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value")
    for r in rows[1:]: print(f, r[ki][:60], r[mi][:45], r[vi])
P
