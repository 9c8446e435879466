#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> <gpus> <script>   -- retries while the pod answers "transient" (nothing charged)
T=$1; G=$2; S=$3
for i in $(seq 1 20); do
  if [ "$G" = "1" ]; then out=$(gpurun --timeout $T -- "bash $S" 2>&1); else out=$(gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=3"; then sleep 90; continue; fi
  break
done
