#!/bin/bash
# Direct DMA of page-locked host buffers: ABI parity tests that use host buffers + the headline bench (e2e line).
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_quant_tc.py -m gpu -x -q > gpurun_out/pytest_s.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_s.log
timeout -s KILL 900 python bench.py > gpurun_out/bench_n1_s.json 2> gpurun_out/bench_n1_s.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/bench_n1_s.json; tail -3 gpurun_out/bench_n1_s.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_n1_s.json'))
print({k:b[k] for k in ('value','ms_per_step','e2e','dtype')}); print(b['roofline']['frac'], b['roofline']['kernel_ms'], b['parity'], b['clocks'])
PY
