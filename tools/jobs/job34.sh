#!/bin/bash
# final multi-GPU C3 points of the session: N=8 and N=4 on the same box
mkdir -p gpurun_out
for n in 8 4 2; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533"
timeout 200 $TR bench.py --gpus $n --steps 10 --warmup 3 --phase-times > gpurun_out/r2_c3_n${n}_final.json 2> gpurun_out/r2_c3_n${n}_final.err; echo "c3 n$n rc=$?"
grep "phases\|self-check\|recall" gpurun_out/r2_c3_n${n}_final.err | head -3; grep "ms/step" gpurun_out/r2_c3_n${n}_final.err | head -2; cut -c1-260 gpurun_out/r2_c3_n${n}_final.json
done
