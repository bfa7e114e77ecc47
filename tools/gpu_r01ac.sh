#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_ac.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_ac.log
timeout -s KILL 300 python tools/bench_configs.py c1 c1big c4 2>&1 | cut -c1-250
timeout -s KILL 200 python tools/qtc_probe.py sq8 10000000 10000 100 4
