#!/bin/bash
# decode-GEMM filter bring-up: parity tests of the new path, then timings of the quantized configs and C5
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_quant_tc.py -x -q > gpurun_out/pytest_qtc_d.log 2>&1; echo "pytest qtc rc=$?"; tail -25 gpurun_out/pytest_qtc_d.log
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q -k "pq_train or opq or kmeans" > gpurun_out/pytest_train_d.log 2>&1; echo "pytest train rc=$?"; tail -5 gpurun_out/pytest_train_d.log
timeout -s KILL 600 python tools/bench_configs.py c2a c2b c3 --small > gpurun_out/configs_small_d.jsonl 2> gpurun_out/configs_small_d.err; echo "configs small rc=$?"; cut -c1-700 gpurun_out/configs_small_d.jsonl; tail -5 gpurun_out/configs_small_d.err
timeout -s KILL 300 python tools/bench_configs.py c5 > gpurun_out/c5_d.jsonl 2> gpurun_out/c5_d.err; echo "c5 rc=$?"; cat gpurun_out/c5_d.jsonl; tail -3 gpurun_out/c5_d.err
