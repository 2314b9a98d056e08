#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3 / transient): $1 = timeout seconds, $2 = command string,
# optional $3 = number of GPUs
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = 1 ]; then out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$CMD" 2>&1); rc=$?
  else out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD" 2>&1); rc=$?; fi
  if echo "$out" | grep -q "status=transient" || [ $rc = 3 ]; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gave up after 40 transient answers"; exit 3
