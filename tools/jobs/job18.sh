#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_build.py tests/test_cpp_mirror.py -x -q -m gpu > gpurun_out/j18_build_tests.log 2>&1; echo "build tests rc=$?"; tail -12 gpurun_out/j18_build_tests.log
timeout 500 python -m pytest tests -q -m gpu > gpurun_out/j18_tests.log 2>&1; echo "all gpu tests rc=$?"; tail -5 gpurun_out/j18_tests.log
timeout 200 python bench.py --steps 5 --warmup 3 --cpu-queries 32 > gpurun_out/j18_c3.json 2> gpurun_out/j18_c3.err; echo "c3 rc=$?"; grep "index K\|recall\|ms/step" gpurun_out/j18_c3.err
timeout 300 python bench.py --config c5 --rows 16000000 --steps 2 --warmup 1 --sweep 64 --no-cpu-baseline > gpurun_out/j18_c5_16m.json 2> gpurun_out/j18_c5_16m.err; echo "c5 16M rc=$?"; grep "trained\|generated\|shard index\|L=\|rror" gpurun_out/j18_c5_16m.err | head
for alg in brute-force partitioned hashed tree-ah; do timeout 120 python tools/ann_benchmark.py --algorithm $alg 2>&1 | grep "json:" ; done > gpurun_out/j18_ann_benchmark.log; cat gpurun_out/j18_ann_benchmark.log | cut -c1-400
