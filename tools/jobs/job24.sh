#!/bin/bash
# mbarrier.try_wait suspend-time hint in tc_scan_kernel: parity tests with it on, then a sweep on C3
mkdir -p gpurun_out
SCANN_TC_WAIT_NS=2000 timeout 300 python -m pytest tests/test_gpu_tcscan.py -q -m gpu > gpurun_out/j24_tests.log 2>&1; echo "tcscan tests (WAIT_NS=2000) rc=$?"; tail -3 gpurun_out/j24_tests.log
export SCANN_TC_DEBUG=1
for w in 0 300 2000 20000; do
  SCANN_TC_WAIT_NS=$w timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --gt-queries 200 > gpurun_out/j24_c3_w$w.json 2> gpurun_out/j24_c3_w$w.err; echo "WAIT_NS=$w rc=$?"
  grep tcscan gpurun_out/j24_c3_w$w.err | tail -1 | cut -c1-120; grep "ms/step" gpurun_out/j24_c3_w$w.err
done
