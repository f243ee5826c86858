#!/bin/bash
# gpurun with retries while the pod answers busy / transient (exit 3 or status=transient): usage: gpurun_retry.sh <timeout> <cmd> [gpus]
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 20); do
  if [ "$G" = "1" ]; then gpurun --timeout $T -- "$CMD" > /tmp/gpurun_try.out 2>&1; else gpurun --gpus $G --timeout $T -- "$CMD" > /tmp/gpurun_try.out 2>&1; fi
  rc=$?
  if grep -q "status=transient\|status=busy" /tmp/gpurun_try.out || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
cat /tmp/gpurun_try.out
