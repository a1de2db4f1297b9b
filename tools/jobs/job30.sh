#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/j30_tests.log 2>&1; echo "tc+parity tests rc=$?"; tail -3 gpurun_out/j30_tests.log
timeout 200 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j30_c2.json 2> gpurun_out/j30_c2.err; echo "c2 rc=$?"; tail -2 gpurun_out/j30_c2.err; python -c "
import json;d=json.loads(open('gpurun_out/j30_c2.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value']);print(json.dumps(d['roofline'])[:500])"
timeout 200 python bench.py --config c2 --sq8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j30_c2_sq8.json 2> gpurun_out/j30_c2_sq8.err; echo "c2 sq8 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/j30_c2_sq8.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'])"
python -c "
import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
