#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_af.json 2> gpurun_out/bench_n2_af.err; echo "bench n2 rc=$?"; head -c 300 gpurun_out/bench_n2_af.json; echo; python -c "
import json
j=json.loads([l for l in open('gpurun_out/bench_n2_af.json') if l.startswith('{')][-1]); print(j['value'], j['e2e'])"
tail -2 gpurun_out/bench_n2_af.err
