#!/bin/bash
mkdir -p gpurun_out
export SCANN_TC_DEBUG=1
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "tests/test_gpu_tcscan.py::test_tc_scan_matches_oracle[150000-128-300-64-500-24-100]" -m gpu -x -q > gpurun_out/j4_san.log 2>&1; echo "san rc=$?" >> gpurun_out/j4_san.log
grep -v "^$" gpurun_out/j4_san.log | head -60
timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/j4_c3.json 2> gpurun_out/j4_c3.err; echo "c3 rc=$?"
grep tcscan gpurun_out/j4_c3.err | tail -3
