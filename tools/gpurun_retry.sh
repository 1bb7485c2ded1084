#!/bin/bash
# local helper (build container): keep asking for a GPU box until the pod has a free slot
# usage: bash tools/gpurun_retry.sh <gpurun args...>
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$OUT" | tail -40
  if echo "$OUT" | grep -q "status=transient\|rc=3\b\|no box\|busy"; then sleep 90; continue; fi
  break
done
