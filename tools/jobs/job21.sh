#!/bin/bash
# 8-GPU round-2 run: C3 scaling point with phase times, C5 (1B x 96 codes, K=65536, L sweep) at size
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
SCANN_TC_DEBUG=1 timeout 200 $TR bench.py --gpus 8 --steps 10 --warmup 3 --phase-times > gpurun_out/r2_c3_n8.json 2> gpurun_out/r2_c3_n8.err; echo "c3 n8 rc=$?"
grep "phases\|self-check\|recall\|ms/step" gpurun_out/r2_c3_n8.err | head -12; grep "tcscan" gpurun_out/r2_c3_n8.err | tail -3; cut -c1-300 gpurun_out/r2_c3_n8.json
timeout 560 $TR bench.py --gpus 8 --config c5 --steps 10 --warmup 3 > gpurun_out/r2_c5_n8.json 2> gpurun_out/r2_c5_n8.err; echo "c5 n8 rc=$?"
grep "checks\|L=\|rror\|trained\|shard plan\|shard index" gpurun_out/r2_c5_n8.err | head -24
nvidia-smi --query-gpu=memory.used --format=csv,noheader | head -2
