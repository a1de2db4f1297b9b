#!/bin/bash
# ncu --set full of tc_score_kernel (DENSE sample launch + FILTER launch, 4096 x 1M x 128 Dot) after the elect.sync change
mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name regex:tc_score_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_tc_score_v3 python tools/bench_bf.py --nq 4096 --reps 1 > gpurun_out/j44_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/j42_ncu.log | cut -c1-200; ls -la gpurun_out/r2_tc_score_v2.ncu-rep
