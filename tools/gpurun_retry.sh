#!/bin/bash
# usage: [GPUS=N] tools/gpurun_retry.sh <timeout_s> '<command>'   — retries while the pod answers busy (rc 3 / transient)
T=$1; shift
G=""; if [ -n "$GPUS" ] && [ "$GPUS" != "1" ]; then G="--gpus $GPUS"; fi
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@"
  rc=$?
  st=$(python -c "import json;print(json.load(open('/root/repo/gpurun_out/.last_call.json')).get('status'))" 2>/dev/null)
  if [ "$st" != "transient" ] && [ "$rc" != "3" ]; then exit $rc; fi
  echo "[retry $i] busy, sleeping 90s"; sleep 90
done
exit 3
