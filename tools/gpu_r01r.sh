#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_quant_tc.py -x -q -k rabitq > gpurun_out/pytest_rabitq_r.log 2>&1; echo "pytest rabitq rc=$?"; tail -25 gpurun_out/pytest_rabitq_r.log
timeout -s KILL 300 python -m pytest tests/test_gpu_quant_tc.py tests/test_gpu_parity.py -x -q > gpurun_out/pytest_all_r.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_all_r.log
timeout -s KILL 600 python tools/bench_configs.py c4 > gpurun_out/c4_r.jsonl 2> gpurun_out/c4_r.err; echo "c4 rc=$?"; cut -c1-900 gpurun_out/c4_r.jsonl; tail -3 gpurun_out/c4_r.err
