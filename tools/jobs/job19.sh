#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_build.py tests/test_kmtree.py tests/test_cpp_mirror.py -q -m gpu > gpurun_out/j19_tests.log 2>&1; echo "build+kmtree tests rc=$?"; tail -15 gpurun_out/j19_tests.log
timeout 300 python bench.py --config c5 --rows 16000000 --steps 2 --warmup 1 --sweep 64 --no-cpu-baseline > gpurun_out/j19_c5_16m.json 2> gpurun_out/j19_c5_16m.err; echo "c5 16M rc=$?"; grep "trained\|generated\|shard index\|L=\|rror" gpurun_out/j19_c5_16m.err | head
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --balance 3 > gpurun_out/j19_c3_bal.json 2> gpurun_out/j19_c3_bal.err; echo "c3 balanced rc=$?"; grep "index K\|recall\|ms/step" gpurun_out/j19_c3_bal.err
