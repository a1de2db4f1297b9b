#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tcscan.py -m gpu -x -q > gpurun_out/j3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j3_tests.log
tail -25 gpurun_out/j3_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/j3_c3.json 2> gpurun_out/j3_c3.err; echo "c3 rc=$?"
tail -4 gpurun_out/j3_c3.err; cat gpurun_out/j3_c3.json | cut -c1-600
