#!/bin/bash
# deep survivor queues + overlapped flush: A/B of the flush threshold on C2 (tc_score) and C3 (tc_scan), parity of both
mkdir -p gpurun_out
for f in 64 320; do
SCANN_TC_WQ_FLUSH=$f timeout 100 python tools/bench_bf.py --nq 4096 --reps 5 > gpurun_out/j41_bf4096_f$f.log 2>&1; echo "bf flush_at=$f rc=$?"; tail -3 gpurun_out/j41_bf4096_f$f.log | cut -c1-150
done
for f in 64 256; do
SCANN_TCS_WQ_FLUSH=$f timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gt-queries 100 > gpurun_out/j41_c3_f$f.json 2> gpurun_out/j41_c3_f$f.err; echo "c3 flush_at=$f rc=$?"
grep "ms/step" gpurun_out/j41_c3_f$f.err; python -c "
import json;d=json.loads(open('gpurun_out/j41_c3_f$f.json').read().strip().splitlines()[-1]);print(d['roofline']['ms_per_launch'], d['recall_at_10'])"
done
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tcscan.py -q -m gpu -x > gpurun_out/j41_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/j41_tests.log
