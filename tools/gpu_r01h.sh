#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_quant_tc.py -x -q > gpurun_out/pytest_qtc_h.log 2>&1; echo "pytest qtc rc=$?"; tail -5 gpurun_out/pytest_qtc_h.log
timeout -s KILL 900 python tools/bench_configs.py c2a c2b c3 > gpurun_out/configs_full_h.jsonl 2> gpurun_out/configs_full_h.err; echo "configs full rc=$?"; cut -c1-330 gpurun_out/configs_full_h.jsonl; tail -5 gpurun_out/configs_full_h.err
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:qtc_kernel -s 1 -c 1 \
    -o gpurun_out/qtc_sq8_full_h -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/qtc_sq8_full_h.log 2>&1; echo "set full rc=$?"
ncu -i gpurun_out/qtc_sq8_full_h.ncu-rep --page raw --csv > gpurun_out/qtc_sq8_full_h_raw.csv 2>/dev/null
ncu -i gpurun_out/qtc_sq8_full_h.ncu-rep --page details > gpurun_out/qtc_sq8_full_h_details.txt 2>/dev/null
grep -E "Duration|L1/TEX Cache Throughput|L2 Cache Throughput|Compute \(SM\)|Issued Warp|No Eligible" gpurun_out/qtc_sq8_full_h_details.txt
