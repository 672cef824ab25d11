#!/bin/bash
# First GPU call of round 2 (profiles/r01_notes.md, last section): everything the last session of round 1 changed without
# being able to run it on a GPU, plus the refreshed headline evidence.  One GPU, about 6 minutes.
#   gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'
set -u
out=gpurun_out/r02_first
mkdir -p "$out"
run() {  # name, timeout, command...
    local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"
}
# 1. the default GPU suite (9-bit digits reach test_config5_full_size only) and the forced-on 9-bit cases
run t_default 400 python -m pytest tests -m gpu -x -q -p no:cacheprovider
SPB_TEST_EXPERIMENTAL=1 run t_nine_bit 200 python -m pytest tests/test_gpu_consolidate.py -k nine_bit -x -q -p no:cacheprovider
# 2. headline line, and the A/Bs of the two defaults that changed (9-bit digits, NUMA binding)
run bench_n1 600 python bench.py
SPB_RADIX9=0 run bench_n1_radix8 200 python bench.py --no-cpu --no-also --no-e2e
SPB_NO_NUMA_BIND=1 run bench_n1_nonuma 300 python bench.py --no-cpu --no-also
run radix9_probe 60 python tools/radix9_probe.py 1e8 3
# 3. launch list of the bench command and one full capture of the 9-bit pass kernel (second launch: the first is a warm-up)
run launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_bench_n1.csv" \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-also
run ncu_radix9 300 ncu --set full --import-source on --clock-control none -k regex:k_radix_pass9 -s 4 -c 1 -o "$out/radix_pass9" \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-also
