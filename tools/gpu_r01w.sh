#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_w.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_w.log
timeout -s KILL 300 python tools/bench_configs.py c1 c4 2>&1 | cut -c1-250
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c1_w.csv python tools/bench_configs.py c1 > /dev/null 2>&1
grep -E "select|flat2|exact_kernel" gpurun_out/launches_c1_w.csv | awk -F'","' '{print $5, $NF}' | tail -4
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5_w.csv python tools/bench_configs.py c5 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_c5_w.csv', errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; kn = h.index('Kernel Name'); mv = h.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    name = r[kn].split('(')[0][:60]
    agg[name][0] += 1; agg[name][1] += v
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]:
    print(f"{t/1e6:10.2f} ms {c:6d}  {k}")
PY
