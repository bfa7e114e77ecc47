#!/bin/bash
# Round-1 GPU evidence pass: parity tests, probe timings, ncu launch list of the bench command,
# one `--set full` capture of the dominant SQ8 scan kernel, DRAM traffic at full size.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/probe.py > gpurun_out/probe.log 2>&1; cat gpurun_out/probe.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:scan_topk_kernel.*SQ8Perm -s 1 -c 1 \
    -o gpurun_out/sq8_full -f python bench.py --steps 1 --warmup 1 --rows 2097152 --queries 2048 --no-cpu-baseline > gpurun_out/sq8_full.log 2>&1
echo "set full rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k regex:scan_topk_kernel.*SQ8Perm -s 1 -c 1 --csv --log-file gpurun_out/sq8_traffic_full.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/sq8_traffic_full.log 2>&1
echo "traffic rc=$?"
ls -la gpurun_out
