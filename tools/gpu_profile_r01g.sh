#!/bin/bash
# BQ through the decode-GEMM filter: parity tests of the quantized tensor-core paths, BQ / C4 / C5 timings.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_quant_tc.py -m gpu -x -q > gpurun_out/pytest_qtc_g.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_qtc_g.log
timeout -s KILL 600 python tools/bench_configs.py bq c4 c5 > gpurun_out/configs_g.jsonl 2> gpurun_out/configs_g.err; echo "configs rc=$?"; cut -c1-420 gpurun_out/configs_g.jsonl; tail -3 gpurun_out/configs_g.err
