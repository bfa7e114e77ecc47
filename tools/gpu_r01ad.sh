#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_quant_tc.py -x -q > gpurun_out/pytest_qtc_ad.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_qtc_ad.log
timeout -s KILL 200 python tools/qtc_probe.py pq 25000000 10000 100 3
