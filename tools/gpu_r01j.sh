#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/bench_configs.py c2b > gpurun_out/c2b_j.jsonl 2>&1; cut -c1-200 gpurun_out/c2b_j.jsonl
timeout -s KILL 300 python tools/bench_configs.py c2b c2a c2b > gpurun_out/c2b_j2.jsonl 2>&1; cut -c1-200 gpurun_out/c2b_j2.jsonl
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum --clock-control none -k regex:qtc_kernel --csv --log-file gpurun_out/launches_c2b_j.csv \
    python tools/bench_configs.py c2b > /dev/null 2>&1; grep -E "qtc_kernel" gpurun_out/launches_c2b_j.csv | cut -d, -f5,13- | head -12
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv
