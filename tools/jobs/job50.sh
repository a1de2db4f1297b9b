#!/bin/bash
# lane-private survivor slots in tc_score_kernel (no queue atomics, one list-slot atomic per lane and flush): C2 + tests
mkdir -p gpurun_out
timeout 60 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j50_c2.json 2> gpurun_out/j50_c2.err; echo "c2 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/j50_c2.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['agreement_with_independent_exact_topk'],d['path_chunks'])"
timeout 60 python bench.py --config c2 --sq8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j50_c2_sq8.json 2> gpurun_out/j50_c2_sq8.err; echo "c2 sq8 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/j50_c2_sq8.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['path_chunks'])"
timeout 120 python -m pytest tests/test_gpu_tc.py tests/test_gpu_r2.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/j50_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/j50_tests.log
