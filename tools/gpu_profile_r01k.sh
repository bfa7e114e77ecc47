#!/bin/bash
# Minima planes: whole GPU suite, shard-of-8 launch list, configs.
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_k.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_k.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2a8_k.csv \
    python tools/bench_configs.py c2a8 > gpurun_out/c2a8_under_ncu_k.log 2>&1; echo "launch list rc=$?"
timeout -s KILL 600 python tools/bench_configs.py c1 c1big c2a8 c2a c2b c4 > gpurun_out/configs_k.jsonl 2> gpurun_out/configs_k.err; echo "configs rc=$?"; cut -c1-300 gpurun_out/configs_k.jsonl; tail -3 gpurun_out/configs_k.err
