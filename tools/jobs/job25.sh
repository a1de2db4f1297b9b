#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tcscan.py -q -m gpu > gpurun_out/j25_tests.log 2>&1; echo "tcscan tests rc=$?"; tail -3 gpurun_out/j25_tests.log
export SCANN_TC_DEBUG=1
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --gt-queries 200 > gpurun_out/j25_c3.json 2> gpurun_out/j25_c3.err; echo "rc=$?"
grep tcscan gpurun_out/j25_c3.err | tail -1 | cut -c1-120; grep "ms/step" gpurun_out/j25_c3.err
