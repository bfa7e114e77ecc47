#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_o.json 2> gpurun_out/bench_n2_o.err; echo "bench n2 rc=$?"; cut -c1-900 gpurun_out/bench_n2_o.json; tail -5 gpurun_out/bench_n2_o.err
