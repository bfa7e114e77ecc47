#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_x.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_x.log
timeout -s KILL 300 python tools/bench_configs.py c1 c4 c5 2>&1 | cut -c1-330
