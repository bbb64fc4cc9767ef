#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_parity.py > gpurun_out/dp_parity.log 2>&1; echo "parity rc=$?"; grep -v "^\*\|OMP" gpurun_out/dp_parity.log | tail -12
for tr in ce nccl; do
ZRB_DP_TRANSPORT=$tr timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/bench_dp2_$tr.json 2> gpurun_out/bench_dp2_$tr.err; echo "bench $tr rc=$?"
python - <<PY
import json
for line in open('gpurun_out/bench_dp2_$tr.json'):
    if line.startswith('{'):
        d=json.loads(line); print('dp2 $tr', round(d['ms_per_step'],3), 'ms', round(d['value']), 'tok/s', d['roofline']['class_ms_per_step'])
PY
grep -v "^\*\|OMP" gpurun_out/bench_dp2_$tr.err | tail -4
done
