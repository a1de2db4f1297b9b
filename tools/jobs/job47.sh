#!/bin/bash
# N = 2 sanity run of the final build under torchrun (NCCL), C3
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 200 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --gt-queries 200 > gpurun_out/j47_c3_n2.json 2> gpurun_out/j47_c3_n2.err; echo "c3 n2 rc=$?"
grep "self-check\|recall" gpurun_out/j47_c3_n2.err | head -3; grep "ms/step" gpurun_out/j47_c3_n2.err | head -2; cut -c1-300 gpurun_out/j47_c3_n2.json
