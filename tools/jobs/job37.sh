#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tcscan.py tests/test_gpu_parity.py tests/test_gpu_split.py -q -m gpu -x > gpurun_out/j37_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/j37_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gt-queries 200 > gpurun_out/j37_c3.json 2> gpurun_out/j37_c3.err; echo "c3 rc=$?"
grep "ms/step\|recall" gpurun_out/j37_c3.err
