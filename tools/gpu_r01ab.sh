#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_flat_tc.py tests/test_gpu_parity.py -x -q > gpurun_out/pytest_gpu_ab.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_ab.log
timeout -s KILL 300 python tools/bench_configs.py c1 c1big 2>&1 | cut -c1-250
