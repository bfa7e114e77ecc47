#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_gpu_al.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/pytest_gpu_al.log
timeout -s KILL 600 python bench.py --no-cpu-baseline > gpurun_out/bench_al.json 2> gpurun_out/bench_al.err; echo "bench rc=$?"; python -c "
import json
j=json.load(open('gpurun_out/bench_al.json')); print(j['value'], j['e2e']['value'], j['roofline']['achieved'], j['roofline']['frac'], j['parity'], j['recall_at_10'])"
