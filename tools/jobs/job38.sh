#!/bin/bash
# TMEM read-bandwidth probe (epilogue bound of tc_score_kernel / tc_scan_kernel)
mkdir -p gpurun_out
timeout 60 tools/tmem_probe.bin > gpurun_out/j38_tmem_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/j38_tmem_probe.log
