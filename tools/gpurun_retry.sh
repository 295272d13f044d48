#!/bin/bash
# usage: gpurun_retry.sh <timeout> <script> [gpus] : retries while the pod has no free slot (nothing is charged for those)
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 12); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "bash $S" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient"; then sleep 100; continue; fi
  echo "$OUT" | tail -40; exit 0
done
echo "gave up: no GPU slot"
