#!/bin/bash
# Local helper: run a command on the GPU box through gpurun, retrying while the pod answers "busy".
# Usage: [GPUS=N] tools/grun.sh TIMEOUT_SECONDS 'command'
T=$1; shift
G=${GPUS:-1}
EXTRA=""
if [ "$G" != "1" ]; then EXTRA="--gpus $G"; fi
for i in $(seq 1 ${RETRIES:-12}); do
  out=$(timeout 3400 /usr/local/graft/bin/gpurun $EXTRA --timeout "$T" -- "$@" 2>&1)
  rc=$?
  if echo "$out" | grep -q "status=ok"; then echo "$out" | tail -4; exit 0; fi
  if echo "$out" | grep -q "status=transient\|retry in a few minutes\|another call"; then sleep 100; continue; fi
  echo "$out" | tail -8; exit $rc
done
echo "grun: gave up after retries"; exit 3
