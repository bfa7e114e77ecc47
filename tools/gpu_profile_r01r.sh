#!/bin/bash
# After the tensor-core PQ-training assignment: whole GPU suite, smoke, C5 A/B at full size (with certificate statistics), C5 launch list and config line.
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r.log
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r.log
timeout -s KILL 600 python tools/pq_assign_ab.py 1000000 768 96 25 gauss > gpurun_out/pq_assign_ab_r.log 2>&1; echo "ab rc=$?"; tail -3 gpurun_out/pq_assign_ab_r.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5_r.csv \
    python tools/c5_dev_only.py > gpurun_out/c5_under_ncu_r.log 2>&1; echo "c5 launch list rc=$?"
timeout -s KILL 300 python tools/bench_configs.py c5 > gpurun_out/configs_c5_r.jsonl 2> gpurun_out/configs_c5_r.err; echo "configs rc=$?"; cut -c1-500 gpurun_out/configs_c5_r.jsonl; tail -2 gpurun_out/configs_c5_r.err
