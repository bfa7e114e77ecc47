#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_gpu_quant_tc.py -x -q -k "sq8_tc_matches_oracle" > gpurun_out/pytest_qtc_l.log 2>&1; echo "pytest sq8 rc=$?"; tail -15 gpurun_out/pytest_qtc_l.log
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
timeout -s KILL 300 python -m pytest tests/test_gpu_quant_tc.py -x -q > gpurun_out/pytest_qtc_l2.log 2>&1; echo "pytest all rc=$?"; tail -8 gpurun_out/pytest_qtc_l2.log
timeout -s KILL 200 python tools/qtc_probe.py sq8 10000000 10000 100 4
timeout -s KILL 200 python tools/qtc_probe.py int4 10000000 10000 100 4
timeout -s KILL 200 python tools/qtc_probe.py pq 25000000 2048 100 4
VECGO_QTC_PAIR=0 timeout -s KILL 200 python tools/qtc_probe.py sq8 10000000 10000 100 3
