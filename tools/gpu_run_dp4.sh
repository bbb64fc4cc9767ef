#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "sparse_embedding or phased" 2>&1 | tail -2
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_parity.py > gpurun_out/dp_parity.log 2>&1; echo "parity rc=$?"; grep -c DP_PARITY_OK gpurun_out/dp_parity.log
for tr in ce nccl; do
ZRB_DP_TRANSPORT=$tr timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2954$((RANDOM%9)) bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/bench_dp4_$tr.json 2> gpurun_out/bench_dp4_$tr.err; echo "bench $tr rc=$?"
python - <<PY
import json
for line in open('gpurun_out/bench_dp4_$tr.json'):
    if line.startswith('{'):
        d=json.loads(line); print('dp4 $tr', round(d['ms_per_step'],3), 'ms', round(d['value']), 'tok/s', d['roofline']['class_ms_per_step'])
PY
grep -v "^\*\|OMP" gpurun_out/bench_dp4_$tr.err | tail -3
done
