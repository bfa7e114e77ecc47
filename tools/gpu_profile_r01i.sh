#!/bin/bash
# Sort-free compaction in the group selection: parity (Flat + quantized filters), per-kernel times of the shard-of-8 step, configs; ncu of pp_dist.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_quant_tc.py tests/test_gpu_flat_tc.py -m gpu -x -q > gpurun_out/pytest_tc_i.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_tc_i.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2a8_i.csv \
    python tools/bench_configs.py c2a8 > gpurun_out/c2a8_under_ncu_i.log 2>&1; echo "launch list rc=$?"
timeout -s KILL 600 python tools/bench_configs.py c1 c1big c2a8 c2a c3 > gpurun_out/configs_i.jsonl 2> gpurun_out/configs_i.err; echo "configs rc=$?"; cut -c1-330 gpurun_out/configs_i.jsonl; tail -3 gpurun_out/configs_i.err
timeout -s KILL 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:pp_dist_kernel -s 10 -c 1 \
    -o gpurun_out/pp_dist_full_i -f python tools/c5_dev_only.py > gpurun_out/pp_dist_full_i.log 2>&1; echo "pp_dist full rc=$?"
ncu -i gpurun_out/pp_dist_full_i.ncu-rep --page details > gpurun_out/pp_dist_full_i_details.txt 2>/dev/null
ncu -i gpurun_out/pp_dist_full_i.ncu-rep --page source --csv --print-source sass > gpurun_out/pp_dist_full_i_src.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
