#!/bin/bash
# Exact stage without the per-CTA row list: parity, per-kernel times of one 1.25M-row shard (8-GPU per-rank work) and of the full C2a.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_quant_tc.py -m gpu -x -q > gpurun_out/pytest_qtc_h.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_qtc_h.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2a8_h.csv \
    python tools/bench_configs.py c2a8 > gpurun_out/c2a8_under_ncu_h.log 2>&1; echo "launch list rc=$?"
timeout -s KILL 600 python tools/bench_configs.py c2a8 c2a c4 > gpurun_out/configs_h.jsonl 2> gpurun_out/configs_h.err; echo "configs rc=$?"; cut -c1-330 gpurun_out/configs_h.jsonl; tail -3 gpurun_out/configs_h.err
