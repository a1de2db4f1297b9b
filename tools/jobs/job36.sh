#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tcscan.py -q -m gpu -x > gpurun_out/j36_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/j36_tests.log
SCANN_TC_DEBUG=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gt-queries 200 > gpurun_out/j36_c3.json 2> gpurun_out/j36_c3.err; echo "c3 rc=$?"
grep tcscan gpurun_out/j36_c3.err | tail -1 | cut -c1-130; grep "ms/step\|recall" gpurun_out/j36_c3.err
