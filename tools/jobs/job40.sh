#!/bin/bash
# overlapped flushes (tc_score_kernel, tc_scan_kernel) + flush after the accumulator hand-over: parity + C2 + C3 timing
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tcscan.py tests/test_gpu_parity.py tests/test_gpu_split.py -q -m gpu -x > gpurun_out/j40_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/j40_tests.log
timeout 200 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j40_c2.json 2> gpurun_out/j40_c2.err; echo "c2 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/j40_c2.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'])"
timeout 200 python bench.py --config c2 --sq8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j40_c2_sq8.json 2> gpurun_out/j40_c2_sq8.err; echo "c2 sq8 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/j40_c2_sq8.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'])"
timeout 100 python tools/bench_bf.py --nq 4096 --reps 5 > gpurun_out/j40_bf4096.log 2>&1; echo "bf rc=$?"; tail -4 gpurun_out/j40_bf4096.log | cut -c1-200
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gt-queries 200 > gpurun_out/j40_c3.json 2> gpurun_out/j40_c3.err; echo "c3 rc=$?"
grep "ms/step\|recall" gpurun_out/j40_c3.err; python -c "
import json;d=json.loads(open('gpurun_out/j40_c3.json').read().strip().splitlines()[-1]);print(d['roofline']['ms_per_launch'], d['roofline']['lut_build_ms_per_launch'], d['stage_ms_per_step'])"
