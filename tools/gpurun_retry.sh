#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <command...>  — retries while the pod answers "transient"/busy (exit 3)
T=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient" || [ $rc -eq 3 ]; then sleep 120; continue; fi
  echo "$out" | tail -40; exit $rc
done
echo "gave up after 20 tries"; exit 3
