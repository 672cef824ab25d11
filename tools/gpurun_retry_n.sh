#!/bin/bash
# gpurun --gpus N with retries on busy
n=$1; t=$2; cmd=$3; log=$4
for attempt in $(seq 1 60); do
    /usr/local/graft/bin/gpurun --gpus "$n" --timeout "$t" -- "$cmd" > "$log" 2>&1
    rc=$?
    if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $attempt)" >> "$log"; exit $rc; fi
    sleep 60
done
exit 3
