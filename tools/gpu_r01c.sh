#!/bin/bash
# HEAD check: GPU parity tests, then a launch list of the C5 training run (where does the time go?)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_c.log
timeout -s KILL 300 python tools/bench_configs.py c5 > gpurun_out/c5_c.jsonl 2> gpurun_out/c5_c.err; echo "c5 rc=$?"; cat gpurun_out/c5_c.jsonl
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5_c.csv \
    python tools/bench_configs.py c5 > gpurun_out/c5_under_ncu.log 2>&1; echo "ncu c5 rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_c5_c.csv', errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; kn = h.index('Kernel Name'); mv = h.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    name = r[kn].split('(')[0][:70]
    agg[name][0] += 1; agg[name][1] += v
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:15]:
    print(f"{t/1e6:10.2f} ms {c:6d}  {k}")
PY
