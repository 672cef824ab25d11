#!/bin/bash
# round 2, call 17 (1 GPU): evidence for the final kernels -- launch list of the bench command, ncu --set full of the dominant
# kernel inside bench.py, and of the dominant kernels of configs 3 and 4
set -u
out=gpurun_out/r02_c17
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_bench_n1.csv" \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-also
run ncu_pass9 400 ncu --set full --import-source on --clock-control none -k regex:k_radix_pass9 -s 6 -c 2 -o "$out/radix_pass9_bulk" \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-also
run ncu_regrid 400 ncu --set full --clock-control none -k regex:"k_merge_numeric|k_merge_count" -s 2 -c 2 -o "$out/merge_regrid" python tools/profile_target.py regrid 1 2
run ncu_rmat 600 ncu --set full --clock-control none -k regex:"k_hash_symbolic|k_hash_numeric" -s 2 -c 2 -o "$out/hash_rmat20" python tools/profile_target.py rmat 20 2
