#!/bin/bash
# k-means / PQ / OPQ training parity after the partition + pick rewrites, launch list and timing of C5.
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kmeans or pq_train or opq or train" > gpurun_out/pytest_kmeans_f.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_kmeans_f.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5_f.csv \
    python tools/c5_dev_only.py > gpurun_out/c5_under_ncu_f.log 2>&1; echo "c5 launch list rc=$?"
timeout -s KILL 300 python tools/bench_configs.py c5 > gpurun_out/configs_c5_f.jsonl 2> gpurun_out/configs_c5_f.err; echo "configs rc=$?"; cut -c1-400 gpurun_out/configs_c5_f.jsonl; tail -3 gpurun_out/configs_c5_f.err
