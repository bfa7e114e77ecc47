#!/bin/bash
# Round-1 evidence pass (final): full GPU test suite, smoke, headline bench (+ reference arm), ncu launch list of the bench
# command, --set full captures of the CTA-pair kernels (SQ8 decode-GEMM, Flat fp16), launch list of the C1 search, every config.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_d.txt 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_d.log
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_d.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_d.log
timeout -s KILL 900 python bench.py > gpurun_out/bench_n1_d.json 2> gpurun_out/bench_n1_d.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_n1_d.json; tail -3 gpurun_out/bench_n1_d.err
timeout -s KILL 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_d.json 2> gpurun_out/bench_ref_d.err; echo "bench ref rc=$?"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_d.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_d.log 2>&1; echo "launch list rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:qtc2_kernel -s 1 -c 1 \
    -o gpurun_out/qtc2_sq8_full_d -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/qtc2_sq8_full_d.log 2>&1; echo "set full rc=$?"
ncu -i gpurun_out/qtc2_sq8_full_d.ncu-rep --page raw --csv > gpurun_out/qtc2_sq8_full_d_raw.csv 2>/dev/null
ncu -i gpurun_out/qtc2_sq8_full_d.ncu-rep --page details > gpurun_out/qtc2_sq8_full_d_details.txt 2>/dev/null
ncu -i gpurun_out/qtc2_sq8_full_d.ncu-rep --page source --csv --print-source sass > gpurun_out/qtc2_sq8_full_d_src.csv 2>/dev/null
timeout -s KILL 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:flat2_kernel -s 1 -c 1 \
    -o gpurun_out/flat2_full_d -f python tools/tc_check.py 1000000 768 2048 10 > gpurun_out/flat2_full_d.log 2>&1; echo "flat2 full rc=$?"
ncu -i gpurun_out/flat2_full_d.ncu-rep --page raw --csv > gpurun_out/flat2_full_d_raw.csv 2>/dev/null
ncu -i gpurun_out/flat2_full_d.ncu-rep --page details > gpurun_out/flat2_full_d_details.txt 2>/dev/null
ncu -i gpurun_out/flat2_full_d.ncu-rep --page source --csv --print-source sass > gpurun_out/flat2_full_d_src.csv 2>/dev/null
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c1_d.csv \
    python tools/bench_configs.py c1 > gpurun_out/c1_under_ncu_d.log 2>&1; echo "c1 launch list rc=$?"
timeout -s KILL 900 python tools/bench_configs.py > gpurun_out/configs_full_d.jsonl 2> gpurun_out/configs_full_d.err; echo "configs rc=$?"; cut -c1-230 gpurun_out/configs_full_d.jsonl; tail -3 gpurun_out/configs_full_d.err
rm -f gpurun_out/*.ncu-rep.tmp
ls -la gpurun_out | tail -8
