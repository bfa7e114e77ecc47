#!/bin/bash
# Tensor-core PQ-training assignment, quick loop: ties A/B, full C5 A/B, C5 launch list.
mkdir -p gpurun_out
timeout -s KILL 300 python tools/pq_assign_ab.py 100000 64 8 5 ties 2>&1 | tail -1
timeout -s KILL 600 python tools/pq_assign_ab.py 1000000 768 96 25 gauss 2>&1 | tail -1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5_p.csv \
    python tools/c5_dev_only.py > gpurun_out/c5_under_ncu_p.log 2>&1; echo "c5 launch list rc=$?"
