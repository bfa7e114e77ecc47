#!/bin/bash
# pp_dist with coalesced loads: training parity + C5 timing / launch list; ncu --set full of the group selection kernel (shard-of-8 step).
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kmeans or pq_train or opq or train" > gpurun_out/pytest_kmeans_j.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_kmeans_j.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5_j.csv \
    python tools/c5_dev_only.py > gpurun_out/c5_under_ncu_j.log 2>&1; echo "c5 launch list rc=$?"
timeout -s KILL 300 python tools/bench_configs.py c5 > gpurun_out/configs_c5_j.jsonl 2> gpurun_out/configs_c5_j.err; echo "configs rc=$?"; cut -c1-400 gpurun_out/configs_c5_j.jsonl
timeout -s KILL 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:tc_select_kernel -s 1 -c 1 \
    -o gpurun_out/select_full_j -f python tools/bench_configs.py c2a8 > gpurun_out/select_full_j.log 2>&1; echo "select full rc=$?"
ncu -i gpurun_out/select_full_j.ncu-rep --page details > gpurun_out/select_full_j_details.txt 2>/dev/null
ncu -i gpurun_out/select_full_j.ncu-rep --page source --csv --print-source sass > gpurun_out/select_full_j_src.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
