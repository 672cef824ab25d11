#!/bin/bash
# round 2, call 44 (1 GPU): pageable transfers with 4 / 8 staging workers (SPB_XFER_WORKERS) -- the C++-API end-to-end program, C-ABI upload/download test
set -u
out=gpurun_out/r02_c44
mkdir -p "$out"
nproc
for w in 4 8; do
  SPB_XFER_WORKERS=$w timeout 200 tests/cpp/_bin/e2e_multiply 100000000 2 1 > "$out/e2e_$w.out" 2> "$out/e2e_$w.err"; echo "workers=$w rc=$?"; tail -n 1 "$out/e2e_$w.out"
done
timeout 200 python -m pytest tests/test_gpu_consolidate.py -x -q -p no:cacheprovider -k "known_answers or fixtures" > "$out/t.out" 2>&1; tail -n 2 "$out/t.out"
