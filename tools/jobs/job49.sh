#!/bin/bash
# sample stride covering the whole row range: brute-force tests (tensor-core path + parity + round-2 cases)
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_tc.py tests/test_gpu_r2.py tests/test_gpu_parity.py -q -m gpu -x -k "bf or sq8 or brute or radius or two_pass or tc_" > gpurun_out/j49_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/j49_tests.log
