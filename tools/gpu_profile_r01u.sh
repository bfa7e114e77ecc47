#!/bin/bash
# Last pass of the round: whole GPU suite on the final library, ncu --set full of the tensor-core assignment kernel.
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_u.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_u.log
timeout -s KILL 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:pqa_kernel -s 2 -c 1 \
    -o gpurun_out/pqa_full_u -f python tools/c5_dev_only.py 262144 > gpurun_out/pqa_full_u.log 2>&1; echo "pqa full rc=$?"
ncu -i gpurun_out/pqa_full_u.ncu-rep --page details > gpurun_out/pqa_full_u_details.txt 2>/dev/null
ncu -i gpurun_out/pqa_full_u.ncu-rep --page raw --csv > gpurun_out/pqa_full_u_raw.csv 2>/dev/null
ncu -i gpurun_out/pqa_full_u.ncu-rep --page source --csv --print-source sass > gpurun_out/pqa_full_u_src.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
