#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name 'regex:^(tc_lut_kernel|tc_scan_kernel)' --launch-skip 4 --launch-count 2 -f -o gpurun_out/r2_tcscan_v3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --gt-queries 10 > gpurun_out/j12_ncu.log 2>&1; echo "ncu rc=$?"
