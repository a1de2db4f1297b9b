#!/bin/bash
# ncu launch list of the C2 step (f32 Dot, 10k x 1M x 128) on the final build
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --config c2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/j46_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/j46_ncu.log | cut -c1-200; wc -l gpurun_out/r2_launches_c2.csv
