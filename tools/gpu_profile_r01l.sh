#!/bin/bash
# Selection with the float prefilter: filter parity tests, shard-of-8 launch list, configs.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_quant_tc.py tests/test_gpu_flat_tc.py -m gpu -x -q > gpurun_out/pytest_tc_l.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_tc_l.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2a8_l.csv \
    python tools/bench_configs.py c2a8 > gpurun_out/c2a8_under_ncu_l.log 2>&1; echo "launch list rc=$?"
timeout -s KILL 600 python tools/bench_configs.py c1 c1big c2a8 c2a > gpurun_out/configs_l.jsonl 2> gpurun_out/configs_l.err; echo "configs rc=$?"; cut -c1-300 gpurun_out/configs_l.jsonl; tail -3 gpurun_out/configs_l.err
