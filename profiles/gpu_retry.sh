#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3 / transient): $1 = timeout seconds, rest = command string
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gave up after 40 transient answers"; exit 3
