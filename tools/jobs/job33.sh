#!/bin/bash
# final single-GPU line of the session + EG=4 check + ncu launch list of the final step
mkdir -p gpurun_out
SCANN_TC_EXP=4 SCANN_TC_DEBUG=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --gt-queries 200 > gpurun_out/j33_c3_eg4.json 2> gpurun_out/j33_c3_eg4.err; echo "EG=4 rc=$?"
grep tcscan gpurun_out/j33_c3_eg4.err | tail -1 | cut -c1-130; grep "ms/step" gpurun_out/j33_c3_eg4.err
timeout 300 python bench.py > gpurun_out/j33_c3.json 2> gpurun_out/j33_c3.err; echo "c3 rc=$?"; grep "index K\|recall\|ms/step" gpurun_out/j33_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:^(tc_|tcwl_|tcs_|wl_|lut16_|merge_|part_|center_|tau_)' -c 4000 --csv --log-file gpurun_out/r2_launches_c3_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --gt-queries 10 > gpurun_out/j33_ncu1.log 2>&1; echo "ncu launches rc=$?"
