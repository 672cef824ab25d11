#!/bin/bash
# round 2, call 43 (1 GPU): C++ layer with parallel pre-faulting of the result vectors -- drop-in tests, then the C++-API end-to-end program
set -u
out=gpurun_out/r02_c43
mkdir -p "$out"
run() { local name=$1 t=$2; shift 2
    ( timeout "$t" "$@" > "$out/$name.out" 2> "$out/$name.err"; echo "rc=$?" >> "$out/$name.err" )
    tail -n 2 "$out/$name.err" | tr '\n' ' '; echo "<- $name"; }
run t_drop 300 python -m pytest tests/test_gpu_dropin.py -x -q -p no:cacheprovider
tail -n 2 "$out/t_drop.out"
run e2e_cpp 300 tests/cpp/_bin/e2e_multiply 100000000 2 1
tail -n 1 "$out/e2e_cpp.out"
