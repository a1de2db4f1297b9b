#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:^(tc_|tcwl_|tcs_|wl_|lut16_|merge_|part_|center_|tau_)' -c 4000 --csv --log-file gpurun_out/r2_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --gt-queries 10 > gpurun_out/j11_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_c3.csv')) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
seq=[(r[ix['Kernel Name']][:70], float(r[ix['Metric Value']].replace(',',''))) for r in rows[1:]]
print(len(seq),"launches")
for n,v in seq[-34:]: print(f"{v/1e6:9.4f} ms  {n}")
PY
