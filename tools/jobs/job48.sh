#!/bin/bash
# adversarial row orders for the two-pass FILTER
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_tc.py -q -m gpu -k "two_pass" > gpurun_out/j48_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/j48_tests.log
