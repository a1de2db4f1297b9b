#!/bin/bash
# 8-GPU round-2 run: C3 scaling point, C4 (100M x 128, SqL2, K=8192) and C5 (1B x 96 codes, K=65536, L sweep) at size
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 200 $TR bench.py --gpus 8 --steps 10 --warmup 3 --phase-times > gpurun_out/r2_c3_n8.json 2> gpurun_out/r2_c3_n8.err; echo "c3 n8 rc=$?"
grep "phases\|self-check\|recall" gpurun_out/r2_c3_n8.err | head -4; cut -c1-200 gpurun_out/r2_c3_n8.json
timeout 330 $TR bench.py --gpus 8 --config c4 --steps 10 --warmup 3 > gpurun_out/r2_c4_n8.json 2> gpurun_out/r2_c4_n8.err; echo "c4 n8 rc=$?"
grep "checks\|L=\|rror\|trained\|shard plan" gpurun_out/r2_c4_n8.err | head -12
timeout 450 $TR bench.py --gpus 8 --config c5 --steps 10 --warmup 3 > gpurun_out/r2_c5_n8.json 2> gpurun_out/r2_c5_n8.err; echo "c5 n8 rc=$?"
grep "checks\|L=\|rror\|trained\|shard plan" gpurun_out/r2_c5_n8.err | head -12
nvidia-smi --query-gpu=memory.used --format=csv,noheader | head -2
