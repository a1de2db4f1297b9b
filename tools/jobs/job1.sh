#!/bin/bash
# round-2 GPU job 1: new tests, bench c3 (ours + reference arm), small single-GPU c4/c5 functional runs
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python -m pytest tests/test_gpu_r2.py tests/test_gpu_parity.py tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/j1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j1_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/j1_c3.json 2> gpurun_out/j1_c3.err; echo "c3 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/j1_ref.json 2> gpurun_out/j1_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --config c4 --n 8000000 --partitions 1024 --steps 3 --warmup 1 --sweep 32,64 > gpurun_out/j1_c4.json 2> gpurun_out/j1_c4.err; echo "c4 rc=$?"
timeout 300 python bench.py --config c5 --n 16000000 --partitions 4096 --latent 4096 --steps 3 --warmup 1 --sweep 16,64 > gpurun_out/j1_c5.json 2> gpurun_out/j1_c5.err; echo "c5 rc=$?"
timeout 200 python bench.py --config c2 --steps 5 --warmup 2 > gpurun_out/j1_c2.json 2> gpurun_out/j1_c2.err; echo "c2 rc=$?"
timeout 100 python bench.py --config c1 --steps 20 --warmup 3 > gpurun_out/j1_c1.json 2> gpurun_out/j1_c1.err; echo "c1 rc=$?"
tail -3 gpurun_out/j1_tests.log
