#!/bin/bash
mkdir -p gpurun_out
export SCANN_TC_DEBUG=1
for c in "600000 128 40 64 20000 8 100"; do
  echo "== $c"; timeout 120 python tools/tc_case.py $c 2>&1 | grep "tcscan\]\|OK\|Error\|error" | head -6
done
