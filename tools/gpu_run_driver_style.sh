#!/bin/bash
# the driver's own commands at N=2: reference arm first, then ours; stdout of each must be exactly one JSON line
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 "$@"; }
run --impl reference > gpurun_out/drv_ref2.out 2> gpurun_out/drv_ref2.err; echo "ref rc=$? lines=$(wc -l < gpurun_out/drv_ref2.out)"
run > gpurun_out/drv_n2.out 2> gpurun_out/drv_n2.err; echo "ours rc=$? lines=$(wc -l < gpurun_out/drv_n2.out)"
python - <<'PY'
import json
for n in ("drv_ref2","drv_n2"):
    t=open(f"gpurun_out/{n}.out").read().strip().splitlines()
    d=json.loads(t[0]); print(n, len(t), d.get("impl"), round(d["value"]), d.get("ms_per_step"), d.get("dp_check"), (d.get("cpu_baseline") or {}).get("cores"), d.get("vs_baseline"))
PY
timeout 420 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -2
