#!/bin/bash
# usage: tools/gpurun_retry.sh <gpus> <timeout> <script>   -- retries while the pod answers "transient" (nothing is charged then)
for i in $(seq 1 12); do
  if [ "$1" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout $2 -- "bash $3" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus $1 --timeout $2 -- "bash $3" 2>&1); fi
  if echo "$out" | grep -q "status=transient"; then echo "[retry $i] busy"; sleep 240; continue; fi
  echo "$out"; exit 0
done
echo "gave up"; exit 3
