#!/bin/bash
# full GPU suite with the concurrency changes, C5-shaped L=256 partition check, final C3 line, ncu launch list + full capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/j22_tests.log 2>&1; echo "all gpu tests rc=$?"; tail -6 gpurun_out/j22_tests.log
timeout 300 python bench.py --config c5 --rows 16000000 --steps 2 --warmup 1 --sweep 64,256 --no-cpu-baseline > gpurun_out/j22_c5_16m.json 2> gpurun_out/j22_c5_16m.err; echo "c5 16M rc=$?"; grep "L=\|rror" gpurun_out/j22_c5_16m.err | head
timeout 300 python bench.py > gpurun_out/j22_c3.json 2> gpurun_out/j22_c3.err; echo "c3 rc=$?"; grep "index K\|recall\|ms/step" gpurun_out/j22_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:^(tc_|tcwl_|tcs_|wl_|lut16_|merge_|part_|center_|tau_)' -c 4000 --csv --log-file gpurun_out/r2_launches_c3_balanced.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --gt-queries 10 > gpurun_out/j22_ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name 'regex:^(tc_lut_kernel|tc_scan_kernel)' --launch-skip 4 --launch-count 2 -f -o gpurun_out/r2_tcscan_v5 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --gt-queries 10 > gpurun_out/j22_ncu2.log 2>&1; echo "ncu full rc=$?"
