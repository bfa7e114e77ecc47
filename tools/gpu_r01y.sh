#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed -k regex:"pp_dist_kernel|pp_pick_kernel|pq_assign_small|part_scatter" -s 40 -c 8 --clock-control none --csv --log-file gpurun_out/c5_kernels_y.csv python tools/bench_configs.py c5 > /dev/null 2>&1
cut -d, -f5,13- gpurun_out/c5_kernels_y.csv | grep -v "^\"ID" | tail -50
