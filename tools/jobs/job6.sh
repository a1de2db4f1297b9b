#!/bin/bash
mkdir -p gpurun_out
export SCANN_TC_DEBUG=1
for c in "60000 128 40 64 500 8 100" "150000 128 300 64 500 24 100" "60000 64 300 32 500 24 100" "60000 128 40 64 50 8 100"; do
  echo "== $c"; timeout 120 python tools/tc_case.py $c 2>&1 | grep "tcscan\]\|OK\|Error\|error" | head -6
done
