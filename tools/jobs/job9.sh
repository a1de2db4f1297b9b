#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tcscan.py -m gpu -q -x > gpurun_out/j9_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/j9_tests.log
export SCANN_TC_DEBUG=1
timeout 200 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/j9_c3.json 2> gpurun_out/j9_c3.err; echo "c3 rc=$?"
grep tcscan gpurun_out/j9_c3.err | tail -1; grep "ms/step" gpurun_out/j9_c3.err
