#!/bin/bash
# k-means++ A/B test on the box, per-kernel launch list of the C5 training, BQ scan baseline at the C4 shape.
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kmeans or pq_train or opq" > gpurun_out/pytest_kmeans_e.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_kmeans_e.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5_e.csv \
    python tools/c5_dev_only.py > gpurun_out/c5_under_ncu_e.log 2>&1; echo "c5 launch list rc=$?"
timeout -s KILL 300 python tools/bench_configs.py bq c5 > gpurun_out/configs_bq_c5_e.jsonl 2> gpurun_out/configs_bq_c5_e.err; echo "configs rc=$?"; cut -c1-400 gpurun_out/configs_bq_c5_e.jsonl; tail -3 gpurun_out/configs_bq_c5_e.err
