#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_quant_tc.py -x -q > gpurun_out/pytest_qtc_f.log 2>&1; echo "pytest qtc rc=$?"; tail -25 gpurun_out/pytest_qtc_f.log
timeout -s KILL 600 python tools/bench_configs.py c2a c2b c3 --small > gpurun_out/configs_small_f.jsonl 2> gpurun_out/configs_small_f.err; echo "configs small rc=$?"; cut -c1-420 gpurun_out/configs_small_f.jsonl; tail -5 gpurun_out/configs_small_f.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_qtc_f.csv \
    python tools/bench_configs.py c2a c2b c3 --small > gpurun_out/qtc_under_ncu_f.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_qtc_f.csv', errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; kn = h.index('Kernel Name'); mv = h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    name = r[kn][:90]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
for k, (c, t) in agg.items():
    if 'qtc' in k or 'tc_select' in k: print(f"{t/1e6/c:10.3f} ms/launch {c:5d}  {k}")
PY
timeout -s KILL 900 python tools/bench_configs.py c2a c2b c3 > gpurun_out/configs_full_f.jsonl 2> gpurun_out/configs_full_f.err; echo "configs full rc=$?"; cut -c1-420 gpurun_out/configs_full_f.jsonl; tail -5 gpurun_out/configs_full_f.err
