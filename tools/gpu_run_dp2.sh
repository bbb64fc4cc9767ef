#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_dp2.json 2> gpurun_out/bench_dp2.err; echo "rc=$?"; cat gpurun_out/bench_dp2.json | head -c 1500; echo; grep -v "^\*\|OMP" gpurun_out/bench_dp2.err | tail -30
