#!/bin/bash
# round 2, call 45 (1 GPU): last GPU seconds -- consolidate / multiply / dense-ops / drop-in suites with the final library
set -u
out=gpurun_out/r02_c45
mkdir -p "$out"
timeout 110 python -m pytest tests/test_gpu_consolidate.py tests/test_gpu_multiply.py tests/test_gpu_dense_ops.py tests/test_gpu_dropin.py -x -q -p no:cacheprovider > "$out/t.out" 2>&1; echo "rc=$?"; tail -n 3 "$out/t.out"
