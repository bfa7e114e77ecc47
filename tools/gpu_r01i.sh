#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_quant_tc.py -x -q > gpurun_out/pytest_qtc_i.log 2>&1; echo "pytest qtc rc=$?"; tail -5 gpurun_out/pytest_qtc_i.log
timeout -s KILL 900 python tools/bench_configs.py c2a c2b c3 > gpurun_out/configs_full_i.jsonl 2> gpurun_out/configs_full_i.err; echo "configs full rc=$?"; cut -c1-330 gpurun_out/configs_full_i.jsonl; tail -5 gpurun_out/configs_full_i.err
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:qtc_kernel -s 1 -c 1 \
    -o gpurun_out/qtc_sq8_full_i -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/qtc_sq8_full_i.log 2>&1; echo "set full rc=$?"
ncu -i gpurun_out/qtc_sq8_full_i.ncu-rep --page raw --csv > gpurun_out/qtc_sq8_full_i_raw.csv 2>/dev/null
ncu -i gpurun_out/qtc_sq8_full_i.ncu-rep --page details > gpurun_out/qtc_sq8_full_i_details.txt 2>/dev/null
grep -E "Duration|L1/TEX Cache Throughput|L2 Cache Throughput|Compute \(SM\)|Issued Warp|No Eligible" gpurun_out/qtc_sq8_full_i_details.txt
grep -E "sm__pipe_tensor_cycles_active_realtime.avg.pct|smsp__inst_executed.sum\"|smsp__issue_active.avg.pct" gpurun_out/qtc_sq8_full_i_raw.csv | head -3
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/qtc_sq8_full_i_raw.csv')))
for h,u,v in zip(rows[0],rows[1],rows[2]):
    if h in ('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed') or 'tensor_cycles_active' in h:
        print(h,u,v)
PY
