#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_quant_tc.py -x -q > gpurun_out/pytest_qtc_t.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_qtc_t.log
timeout -s KILL 200 python tools/qtc_probe.py sq8 10000000 10000 100 4
timeout -s KILL 200 python tools/qtc_probe.py int4 10000000 10000 100 3
timeout -s KILL 300 python tools/bench_configs.py c4 2>&1 | cut -c1-330
