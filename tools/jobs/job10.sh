#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tcscan.py -m gpu -q -x > gpurun_out/j10_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/j10_tests.log
export SCANN_TC_DEBUG=1
timeout 200 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/j10_c3.json 2> gpurun_out/j10_c3.err; echo "c3 rc=$?"
grep tcscan gpurun_out/j10_c3.err | tail -1; grep "ms/step" gpurun_out/j10_c3.err
unset SCANN_TC_DEBUG
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:tc_scan_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_tcscan_v2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --gt-queries 10 > gpurun_out/j10_ncu.log 2>&1; echo "ncu rc=$?"
