#!/bin/bash
# Tensor-core PQ-training assignment: A/B against the exact assignment (three data kinds + the C5 shape), training parity tests, C5 launch list.
mkdir -p gpurun_out
for kind in gauss ties scaled; do timeout -s KILL 300 python tools/pq_assign_ab.py 100000 64 8 5 $kind 2>&1 | tail -2; done
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kmeans or pq_train or opq or train" > gpurun_out/pytest_kmeans_n.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_kmeans_n.log
timeout -s KILL 600 python tools/pq_assign_ab.py 1000000 768 96 25 gauss 2>&1 | tail -2
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5_n.csv \
    python tools/c5_dev_only.py > gpurun_out/c5_under_ncu_n.log 2>&1; echo "c5 launch list rc=$?"
timeout -s KILL 300 python tools/bench_configs.py c5 > gpurun_out/configs_c5_n.jsonl 2> gpurun_out/configs_c5_n.err; echo "configs rc=$?"; cut -c1-300 gpurun_out/configs_c5_n.jsonl; tail -2 gpurun_out/configs_c5_n.err
