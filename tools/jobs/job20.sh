#!/bin/bash
# balanced C3 index: how many closest leaves (T) should the register-LUT kernel take before the tensor-core pass?
mkdir -p gpurun_out
export SCANN_TC_DEBUG=1
for t in 2 3 4 6; do
  SCANN_TC_RANKS=$t timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --gt-queries 200 --balance 3 > gpurun_out/j20_c3_T$t.json 2> gpurun_out/j20_c3_T$t.err; echo "T=$t rc=$?"
  grep tcscan gpurun_out/j20_c3_T$t.err | tail -1; grep "ms/step\|recall" gpurun_out/j20_c3_T$t.err
done
