#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:pqa_kernel -s 2 -c 1 \
    -o gpurun_out/pqa_full_q -f python tools/c5_dev_only.py 262144 > gpurun_out/pqa_full_q.log 2>&1; echo "pqa full rc=$?"
ncu -i gpurun_out/pqa_full_q.ncu-rep --page details > gpurun_out/pqa_full_q_details.txt 2>/dev/null
ncu -i gpurun_out/pqa_full_q.ncu-rep --page source --csv --print-source sass > gpurun_out/pqa_full_q_src.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
