#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for ov in 0 1; do
ZRB_DP_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$ov bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_dp8_ov$ov.json 2> gpurun_out/bench_dp8_ov$ov.err; echo "rc=$?"
python - <<PY
import json
for line in open('gpurun_out/bench_dp8_ov$ov.json'):
    if line.startswith('{'):
        d=json.loads(line); print('dp8 overlap=$ov', round(d['ms_per_step'],3), 'ms', round(d['value']), 'tok/s', d['roofline']['class_ms_per_step'])
PY
grep -v "^\*\|OMP" gpurun_out/bench_dp8_ov$ov.err | tail -3
done
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 2>&1 | grep -iE "nvls|channels|Connected|algo" | head -12 > gpurun_out/nccl_info.txt; cat gpurun_out/nccl_info.txt | head -12
