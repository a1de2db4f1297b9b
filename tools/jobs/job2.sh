#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/tcscan_probe.bin > gpurun_out/j2_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/j2_probe.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j2_tests.log
cat gpurun_out/j2_probe.log; tail -15 gpurun_out/j2_tests.log
