#!/bin/bash
# token prefetch pipeline: GPU test, then C3 at N=2 with and without it (phase times + self-check)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_split.py -q -m gpu > gpurun_out/j27_tests.log 2>&1; echo "split tests rc=$?"; tail -4 gpurun_out/j27_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 200 $TR bench.py --gpus 2 --steps 10 --warmup 3 --phase-times --no-prefetch > gpurun_out/j27_c3_n2_nopf.json 2> gpurun_out/j27_c3_n2_nopf.err; echo "n2 no prefetch rc=$?"
grep "phases\|self-check\|recall\|ms/step" gpurun_out/j27_c3_n2_nopf.err | head -6
timeout 200 $TR bench.py --gpus 2 --steps 10 --warmup 3 --phase-times > gpurun_out/j27_c3_n2.json 2> gpurun_out/j27_c3_n2.err; echo "n2 prefetch rc=$?"
grep "phases\|self-check\|recall\|ms/step" gpurun_out/j27_c3_n2.err | head -6
