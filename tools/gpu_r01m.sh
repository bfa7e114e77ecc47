#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:qtc2_kernel -s 1 -c 1 \
    -o gpurun_out/qtc2_sq8_full_m -f python tools/qtc_probe.py sq8 10000000 10000 100 2 > gpurun_out/qtc2_sq8_full_m.log 2>&1; echo "set full rc=$?"
ncu -i gpurun_out/qtc2_sq8_full_m.ncu-rep --page raw --csv > gpurun_out/qtc2_sq8_full_m_raw.csv 2>/dev/null
ncu -i gpurun_out/qtc2_sq8_full_m.ncu-rep --page details > gpurun_out/qtc2_sq8_full_m_details.txt 2>/dev/null
ncu -i gpurun_out/qtc2_sq8_full_m.ncu-rep --page source --csv --print-source sass > gpurun_out/qtc2_sq8_full_m_src.csv 2>/dev/null
grep -E "Duration|L1/TEX Cache Throughput|L2 Cache Throughput|Compute \(SM\)|Issued Warp|No Eligible|DRAM Throughput" gpurun_out/qtc2_sq8_full_m_details.txt
