#!/bin/bash
# ONE 8-GPU box: data-parallel bench at N = 8 and 4 through both transports, then the 10-model Large ensemble sharded
# one model per GPU (BASELINE configs[3], [4]).  Strict per-step timeouts: a hang must not eat the box.
mkdir -p gpurun_out
for N in 8 4; do
for tr in nccl ce; do
  ZRB_DP_TRANSPORT=$tr timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
     bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_dp${N}_$tr.json 2> gpurun_out/bench_dp${N}_$tr.err
  echo "dp$N $tr rc=$?"
done; done
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 \
   tools/ensemble_eval.py --recipe large --ensemble_num 10 --epochs 4 --json gpurun_out/ensemble_large10_8gpu.json > gpurun_out/ensemble_large10_8gpu.log 2>&1; echo "ensemble rc=$?"; grep -E "averaged models|model [0-9]+:" gpurun_out/ensemble_large10_8gpu.log | tail -8
python - <<'PY'
import json
for N in (8,4):
  for tr in ("nccl","ce"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/bench_dp{N}_{tr}.json") if l.startswith("{")][0])
        print(N, tr, "ms/step", round(d["ms_per_step"],4), "tok/s", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],4), d.get("dp_check",{}).get("replicas_identical"), d["roofline"]["class_ms_per_step"])
    except Exception as e: print(N, tr, "failed", e)
PY
