#!/bin/bash
# round 2, call 31 (1 GPU): compute-sanitizer (memcheck, racecheck) over the kernel families incl. the new reduce / walk / bitmap-walk paths
set -u
out=gpurun_out/r02_c31
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run plain 300 python tools/sanitize_target.py
tail -n 2 "$out/plain.out"
run memcheck 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_target.py
grep -E "ERROR SUMMARY|Invalid|sanitize target" "$out/memcheck.out" | head -8
run racecheck 1500 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_target.py
grep -E "RACECHECK SUMMARY|hazard|sanitize target" "$out/racecheck.out" | head -12
run t_cons 900 python -m pytest tests/test_gpu_consolidate.py -x -q -p no:cacheprovider -k "row_passes"
tail -n 3 "$out/t_cons.out"
