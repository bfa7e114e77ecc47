#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_full_size.py -x -q --durations=5 > gpurun_out/pytest_full_ah.log 2>&1; echo "pytest rc=$?"; tail -16 gpurun_out/pytest_full_ah.log
