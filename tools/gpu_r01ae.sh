#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_n4_ae.json 2> gpurun_out/bench_n4_ae.err; echo "bench n4 rc=$?"; cut -c1-400 gpurun_out/bench_n4_ae.json; tail -3 gpurun_out/bench_n4_ae.err
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_ae.json 2> gpurun_out/bench_n2_ae.err; echo "bench n2 rc=$?"; cut -c1-200 gpurun_out/bench_n2_ae.json
