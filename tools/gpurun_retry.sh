#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged).  usage: tools/gpurun_retry.sh <timeout_s> '<command>' [log]
t=$1; cmd=$2; log=${3:-/tmp/gpurun_last.log}
for attempt in $(seq 1 40); do
    /usr/local/graft/bin/gpurun --timeout "$t" -- "$cmd" > "$log" 2>&1
    rc=$?
    if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $attempt)" >> "$log"; exit $rc; fi
    sleep 60
done
echo "gpurun: gave up after 40 busy answers" >> "$log"; exit 3
