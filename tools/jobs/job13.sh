#!/bin/bash
mkdir -p gpurun_out
export SCANN_TC_DEBUG=1
for e in 2 4; do
  SCANN_TC_EXP=$e timeout 200 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --gt-queries 100 > gpurun_out/j13_c3_$e.json 2> gpurun_out/j13_c3_$e.err; echo "c3 EG=$e rc=$?"
  grep tcscan gpurun_out/j13_c3_$e.err | tail -1; grep "ms/step\|recall" gpurun_out/j13_c3_$e.err
done
