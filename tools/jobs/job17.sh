#!/bin/bash
# round 2, session 2: native index build (build_index.cu) tests, full GPU suite, C3 with the in-library trainer
# (plain and balanced), C5-shaped assignment speed at K = 65,536
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_build.py tests/test_cpp_mirror.py -x -q -m gpu > gpurun_out/j17_build_tests.log 2>&1; echo "build tests rc=$?"; tail -15 gpurun_out/j17_build_tests.log
timeout 500 python -m pytest tests -x -q -m gpu > gpurun_out/j17_tests.log 2>&1; echo "all gpu tests rc=$?"; tail -5 gpurun_out/j17_tests.log
timeout 200 python bench.py --steps 5 --warmup 3 --cpu-queries 32 > gpurun_out/j17_c3.json 2> gpurun_out/j17_c3.err; echo "c3 rc=$?"; grep "index K\|recall\|ms/step" gpurun_out/j17_c3.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --balance 3 > gpurun_out/j17_c3_bal.json 2> gpurun_out/j17_c3_bal.err; echo "c3 balanced rc=$?"; grep "index K\|recall\|ms/step" gpurun_out/j17_c3_bal.err
timeout 300 python bench.py --config c5 --rows 16000000 --steps 2 --warmup 1 --sweep 64 --no-cpu-baseline > gpurun_out/j17_c5_16m.json 2> gpurun_out/j17_c5_16m.err; echo "c5 16M rc=$?"; grep "trained\|generated\|shard index\|L=\|rror" gpurun_out/j17_c5_16m.err | head
