#!/bin/bash
# usage: tools/gpurun_retry.sh [--gpus N] TIMEOUT 'command'   - retries while the pod answers "transient"/busy
GPUS=""
if [ "$1" == "--gpus" ]; then GPUS="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 12); do
  OUT=$(/usr/local/graft/bin/gpurun $GPUS --timeout $T -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient\|no box or slot"; then sleep 120; continue; fi
  echo "$OUT"; exit 0
done
echo "$OUT"; echo "gave up after retries"
