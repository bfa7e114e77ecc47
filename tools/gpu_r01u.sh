#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_quant_tc.py -x -q > gpurun_out/pytest_qtc_u.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_qtc_u.log
timeout -s KILL 200 python tools/qtc_probe.py sq8 10000000 10000 100 4
timeout -s KILL 200 python tools/qtc_probe.py int4 10000000 10000 100 3
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_probe_u.csv python tools/qtc_probe.py sq8 10000000 10000 100 2 > /dev/null 2>&1
grep -E "exact_kernel|select_kernel|qtc2_kernel" gpurun_out/launches_probe_u.csv | awk -F'","' '{print $5, $NF}' | tail -6
