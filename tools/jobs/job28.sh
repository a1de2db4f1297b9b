#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/j28_ref.json 2> gpurun_out/j28_ref.err ) 2>&1 | grep real; echo "ref rc=$?"; tail -3 gpurun_out/j28_ref.err; cut -c1-600 gpurun_out/j28_ref.json
python - <<'PY'
import subprocess,sys
# which .so files did the reference arm load?
import json
PY
for t in 1; do SCANN_TC_RANKS=$t SCANN_TC_DEBUG=1 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --gt-queries 200 > gpurun_out/j28_c3_T$t.json 2> gpurun_out/j28_c3_T$t.err; echo "T=$t rc=$?"; grep tcscan gpurun_out/j28_c3_T$t.err | tail -1 | cut -c1-200; grep "ms/step\|recall" gpurun_out/j28_c3_T$t.err; done
