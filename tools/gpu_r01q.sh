#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_flat_tc.py -x -q > gpurun_out/pytest_flat_q.log 2>&1; echo "pytest flat rc=$?"; tail -12 gpurun_out/pytest_flat_q.log
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity_q.log 2>&1; echo "pytest parity rc=$?"; tail -4 gpurun_out/pytest_parity_q.log
timeout -s KILL 600 python tools/bench_configs.py c1 c1big > gpurun_out/configs_flat_q.jsonl 2> gpurun_out/configs_flat_q.err; echo "configs rc=$?"; cut -c1-420 gpurun_out/configs_flat_q.jsonl; tail -3 gpurun_out/configs_flat_q.err
VECGO_FLAT_PAIR=0 timeout -s KILL 600 python tools/bench_configs.py c1big > gpurun_out/configs_flat_q0.jsonl 2>&1; cut -c1-250 gpurun_out/configs_flat_q0.jsonl
