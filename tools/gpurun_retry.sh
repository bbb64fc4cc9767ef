#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <gpus> <script>   -- retries while the pod has no free slot
T=$1; G=$2; S=$3
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- bash $S 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 75; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
