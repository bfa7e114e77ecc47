#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_ag.json 2> gpurun_out/bench_n2_ag.err; echo "bench n2 rc=$?"; wc -l gpurun_out/bench_n2_ag.json; head -c 120 gpurun_out/bench_n2_ag.json; echo; grep -c "NCCL version" gpurun_out/bench_n2_ag.err
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2_ag.json 2> /dev/null; echo "ref rc=$?"; wc -l gpurun_out/bench_ref_n2_ag.json; head -c 150 gpurun_out/bench_ref_n2_ag.json
