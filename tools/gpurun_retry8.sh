#!/bin/bash
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --gpus 8 --timeout 1000 -- 'bash tools/jobs/job34.sh' 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  echo "$out" | tail -60; exit $rc
done
echo "gave up"; exit 3
