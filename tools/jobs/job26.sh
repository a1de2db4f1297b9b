#!/bin/bash
mkdir -p gpurun_out
export SCANN_TC_DEBUG=1
for w in 0 40 200; do
  SCANN_TC_WAIT_NS=$w timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --gt-queries 200 > gpurun_out/j26_c3_w$w.json 2> gpurun_out/j26_c3_w$w.err; echo "WAIT_NS=$w rc=$?"
  grep tcscan gpurun_out/j26_c3_w$w.err | tail -1 | cut -c1-120; grep "ms/step" gpurun_out/j26_c3_w$w.err
done
