#!/bin/bash
# strict alternation of the two tc_scan pipelines: parity tests with it on, then A/B on C3
mkdir -p gpurun_out
SCANN_TC_ALT=1 timeout 300 python -m pytest tests/test_gpu_tcscan.py tests/test_gpu_split.py -q -m gpu > gpurun_out/j23_tests.log 2>&1; echo "tcscan tests (ALT=1) rc=$?"; tail -4 gpurun_out/j23_tests.log
export SCANN_TC_DEBUG=1
for alt in 0 1; do
  SCANN_TC_ALT=$alt timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --gt-queries 200 > gpurun_out/j23_c3_alt$alt.json 2> gpurun_out/j23_c3_alt$alt.err; echo "ALT=$alt rc=$?"
  grep tcscan gpurun_out/j23_c3_alt$alt.err | tail -1; grep "ms/step\|recall" gpurun_out/j23_c3_alt$alt.err
done
