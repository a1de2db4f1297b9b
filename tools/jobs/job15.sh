#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 2 --cpu-queries 32 > gpurun_out/j15_c3.json 2> gpurun_out/j15_c3.err; echo "c3 rc=$?"; tail -2 gpurun_out/j15_c3.err
python -c "
import json;d=json.load(open('gpurun_out/j15_c3.json'));print(d['value'],d['ms_per_step'],d['recall_at_10']);print(json.dumps(d['roofline'])[:900]);print(d['scan_path'],d.get('cpu_baseline'))"
timeout 300 python bench.py --config c4 --rows 4000000 --partitions 512 --steps 2 --warmup 1 --sweep 32,64 > gpurun_out/j15_c4.json 2> gpurun_out/j15_c4.err; echo "c4 rc=$?"; tail -2 gpurun_out/j15_c4.err
python -c "
import json;d=json.load(open('gpurun_out/j15_c4.json'));print(d['value'],d['ms_per_step'],d['recall_at_10']);print(json.dumps(d['roofline'])[:600]);print(d['scan_path'])"
