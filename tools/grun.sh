#!/bin/bash
# Local helper: run a command on the GPU box through gpurun, retrying while the pod answers "busy".
# Usage: tools/grun.sh TIMEOUT_SECONDS 'command'
T=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10; do
  out=$(timeout 3400 /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  rc=$?
  if echo "$out" | grep -q "status=ok"; then echo "$out" | tail -3; exit 0; fi
  if echo "$out" | grep -q "status=transient\|retry in a few minutes"; then sleep 90; continue; fi
  echo "$out" | tail -8; exit $rc
done
echo "grun: gave up after retries"; exit 3
