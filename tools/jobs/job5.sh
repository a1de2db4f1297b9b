#!/bin/bash
mkdir -p gpurun_out
export SCANN_TC_DEBUG=1
for t in "30000-32-12-16-300-4-50" "120000-96-60-48-700-16-100" "150000-128-300-64-500-24-100" "60000-64-8-32-1000-8-100" "20000-96-40-48-37-6-20"; do
  timeout 200 python -m pytest "tests/test_gpu_tcscan.py::test_tc_scan_matches_oracle[$t]" -m gpu -x -q > gpurun_out/j5_t_$t.log 2>&1; echo "$t rc=$?"
  grep "tcscan\]\|Error\|error" gpurun_out/j5_t_$t.log | head -5
done
timeout 200 python -m pytest tests/test_gpu_tcscan.py -m gpu -q -k "same_results or flagged" > gpurun_out/j5_t_other.log 2>&1; echo "other rc=$?"; tail -5 gpurun_out/j5_t_other.log
for T in 1 2 4; do
  SCANN_TC_RANKS=$T timeout 200 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/j5_c3_T$T.json 2> gpurun_out/j5_c3_T$T.err; echo "c3 T=$T rc=$?"
  grep tcscan gpurun_out/j5_c3_T$T.err | tail -1; grep "ms/step" gpurun_out/j5_c3_T$T.err
done
