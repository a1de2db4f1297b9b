#!/bin/bash
# tc_score_kernel with the compact FILTER epilogue (lane-local pushes, per-tile flush check, single-copy flush): parity + C2 timing
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_r2.py tests/test_gpu_build.py -q -m gpu -x > gpurun_out/j43_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/j43_tests.log
timeout 200 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j43_c2.json 2> gpurun_out/j43_c2.err; echo "c2 rc=$?"; tail -2 gpurun_out/j43_c2.err; python -c "
import json;d=json.loads(open('gpurun_out/j43_c2.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value']);print(json.dumps(d['roofline'])[:300])"
timeout 200 python bench.py --config c2 --sq8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j43_c2_sq8.json 2> gpurun_out/j43_c2_sq8.err; echo "c2 sq8 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/j43_c2_sq8.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'])"
timeout 100 python tools/bench_bf.py --nq 4096 --reps 5 > gpurun_out/j43_bf4096.log 2>&1; echo "bf rc=$?"; tail -4 gpurun_out/j43_bf4096.log | cut -c1-400
