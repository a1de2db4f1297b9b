#!/bin/bash
# two-pass FILTER (prefix pass -> tightened bound -> rest): C2 timing first, then the whole GPU suite on this build, then C3
mkdir -p gpurun_out
timeout 200 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j45_c2.json 2> gpurun_out/j45_c2.err; echo "c2 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/j45_c2.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['agreement_with_independent_exact_topk'],d['path_chunks'])"
timeout 200 python bench.py --config c2 --sq8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j45_c2_sq8.json 2> gpurun_out/j45_c2_sq8.err; echo "c2 sq8 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/j45_c2_sq8.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['path_chunks'])"
timeout 100 python bench.py --config c1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j45_c1.json 2> gpurun_out/j45_c1.err; echo "c1 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/j45_c1.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'])"
timeout 500 python -m pytest tests -q -m gpu -x > gpurun_out/j45_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/j45_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gt-queries 200 > gpurun_out/j45_c3.json 2> gpurun_out/j45_c3.err; echo "c3 rc=$?"; grep "ms/step\|recall" gpurun_out/j45_c3.err
