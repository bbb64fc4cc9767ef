#!/bin/bash
# N=2 bench only (default transport), no CPU baseline
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_dp2_ce.json 2> gpurun_out/bench_dp2_ce.err; echo "rc=$?"
python - <<PY
import json
for line in open('gpurun_out/bench_dp2_ce.json'):
    if line.startswith('{'):
        d=json.loads(line); print('dp2', d['config'].get('dp_transport'), round(d['ms_per_step'],3), 'ms', round(d['value']), 'tok/s', d['roofline']['class_ms_per_step'])
PY
tail -2 gpurun_out/bench_dp2_ce.err
