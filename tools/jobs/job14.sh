#!/bin/bash
mkdir -p gpurun_out
export SCANN_TC_DEBUG=0
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 2 --steps 5 --warmup 3 --phase-times --check-queries 64 > gpurun_out/j14_c3_n2.json 2> gpurun_out/j14_c3_n2.err; echo "c3 n2 rc=$?"
grep "phases\|ms/step\|self-check\|recall" gpurun_out/j14_c3_n2.err | head
timeout 400 $TR bench.py --gpus 2 --config c4 --rows 16000000 --partitions 2048 --steps 3 --warmup 2 --sweep 32,64 > gpurun_out/j14_c4_n2.json 2> gpurun_out/j14_c4_n2.err; echo "c4 n2 rc=$?"
grep "checks\|L=\|Error\|error" gpurun_out/j14_c4_n2.err | head
timeout 400 $TR bench.py --gpus 2 --config c5 --rows 32000000 --partitions 8192 --latent 8192 --steps 3 --warmup 2 --sweep 16,64 > gpurun_out/j14_c5_n2.json 2> gpurun_out/j14_c5_n2.err; echo "c5 n2 rc=$?"
grep "checks\|L=\|Error\|error" gpurun_out/j14_c5_n2.err | head
