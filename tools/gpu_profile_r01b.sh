#!/bin/bash
# Round-1 evidence pass (second half of the round): headline bench, per-config bench, ncu launch lists and
# --set full captures of the dominant kernels, full-size DRAM traffic of the SQ8 scan.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_b.txt 2>&1
timeout -s KILL 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_n1.json
timeout -s KILL 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout -s KILL 900 python tools/bench_configs.py > gpurun_out/configs_full.jsonl 2> gpurun_out/configs_full.err; echo "configs rc=$?"; cut -c1-300 gpurun_out/configs_full.jsonl
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_b.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_b.log 2>&1; echo "launch list rc=$?"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:scan_topk_kernel.*SQ8Perm -s 1 -c 1 \
    -o gpurun_out/sq8_full_b -f python bench.py --steps 1 --warmup 1 --rows 2097152 --queries 2048 --no-cpu-baseline > gpurun_out/sq8_full_b.log 2>&1; echo "sq8 full rc=$?"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:flat_tc_kernel -s 1 -c 1 \
    -o gpurun_out/flat_tc_full_b -f python tools/tc_check.py 1000000 768 2048 10 > gpurun_out/flat_tc_full_b.log 2>&1; echo "flat full rc=$?"
timeout -s KILL 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k regex:scan_topk_kernel.*SQ8Perm -s 1 -c 1 --csv --log-file gpurun_out/sq8_traffic_full_b.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/sq8_traffic_full_b.log 2>&1; echo "traffic rc=$?"
ls -la gpurun_out | tail -20
