#!/bin/bash
# Round-1 evidence pass for the tensor-core quantized scans: full GPU test suite, headline bench (+ reference arm),
# ncu launch list of the bench command, --set full capture and DRAM traffic of the GEMM kernel at full size, every config.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_c.txt 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_c2.log
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_c.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_c.log
timeout -s KILL 900 python bench.py > gpurun_out/bench_n1_c.json 2> gpurun_out/bench_n1_c.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_n1_c.json; tail -3 gpurun_out/bench_n1_c.err
timeout -s KILL 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_c.json 2> gpurun_out/bench_ref_c.err; echo "bench ref rc=$?"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_c.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_c.log 2>&1; echo "launch list rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:qtc2_kernel -s 2 -c 1 \
    -o gpurun_out/qtc2_sq8_full_c -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/qtc2_sq8_full_c.log 2>&1; echo "set full rc=$?"
ncu -i gpurun_out/qtc2_sq8_full_c.ncu-rep --page raw --csv > gpurun_out/qtc2_sq8_full_c_raw.csv 2>/dev/null
ncu -i gpurun_out/qtc2_sq8_full_c.ncu-rep --page details > gpurun_out/qtc2_sq8_full_c_details.txt 2>/dev/null
timeout -s KILL 900 python tools/bench_configs.py > gpurun_out/configs_full_c.jsonl 2> gpurun_out/configs_full_c.err; echo "configs rc=$?"; cut -c1-260 gpurun_out/configs_full_c.jsonl
ls -la gpurun_out | tail -20
ncu -i gpurun_out/qtc2_sq8_full_c.ncu-rep --page source --csv --print-source sass > gpurun_out/qtc2_sq8_full_c_src.csv 2>/dev/null
