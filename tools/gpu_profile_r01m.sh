#!/bin/bash
# Round-1 evidence pass (final state): full GPU suite, smoke, headline bench + reference arm, ncu launch list of the bench command,
# launch lists of the C1 and C5 paths, every config on one GPU.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_m.txt 2>&1
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_m.log
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_m.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_m.log
timeout -s KILL 900 python bench.py > gpurun_out/bench_n1_m.json 2> gpurun_out/bench_n1_m.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_n1_m.json; tail -3 gpurun_out/bench_n1_m.err
timeout -s KILL 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_m.json 2> gpurun_out/bench_ref_m.err; echo "bench ref rc=$?"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_m.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_m.log 2>&1; echo "launch list rc=$?"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c1_m.csv \
    python tools/bench_configs.py c1 > gpurun_out/c1_under_ncu_m.log 2>&1; echo "c1 launch list rc=$?"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5_m.csv \
    python tools/c5_dev_only.py > gpurun_out/c5_under_ncu_m.log 2>&1; echo "c5 launch list rc=$?"
timeout -s KILL 900 python tools/bench_configs.py c1 c1big c2a c2b c3 c4 bq c5 c2a8 > gpurun_out/configs_full_m.jsonl 2> gpurun_out/configs_full_m.err; echo "configs rc=$?"; cut -c1-230 gpurun_out/configs_full_m.jsonl; tail -3 gpurun_out/configs_full_m.err
ls -la gpurun_out | tail -5
