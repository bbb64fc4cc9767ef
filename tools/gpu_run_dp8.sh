#!/bin/bash
mkdir -p gpurun_out
ZRB_DP_TRANSPORT=ce timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/bench_dp8_ce.json 2> gpurun_out/bench_dp8_ce.err; echo "bench ce rc=$?"
python - <<PY
import json
for line in open('gpurun_out/bench_dp8_ce.json'):
    if line.startswith('{'):
        d=json.loads(line); print('dp8 ce', round(d['ms_per_step'],3), 'ms', round(d['value']), 'tok/s', d['roofline']['class_ms_per_step'])
PY
grep -v "^\*\|OMP" gpurun_out/bench_dp8_ce.err | tail -4
