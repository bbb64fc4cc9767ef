#!/bin/bash
# usage: tools/gpu_run_dp.sh N [steps]   -- multi-GPU parity tests (N >= 2) + bench at N ranks through every transport
N=${1:-2}; STEPS=${2:-100}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 420 python -m pytest tests/test_gpu_multi.py -q -x -s > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?" >> gpurun_out/pytest_multi.log; grep -E "^DP |passed|failed|Error|error" gpurun_out/pytest_multi.log | cut -c1-600 | tail -12
fi
for tr in ce nccl; do
  ZRB_DP_TRANSPORT=$tr timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps $STEPS --warmup 10 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_dp${N}_$tr.json 2> gpurun_out/bench_dp${N}_$tr.err
  echo "dp$N $tr rc=$?"; tail -2 gpurun_out/bench_dp${N}_$tr.err | cut -c1-300
done
python - <<PY
import json
for tr in ("ce","nccl"):
    try:
        d=json.loads([l for l in open("gpurun_out/bench_dp${N}_%s.json" % tr) if l.startswith("{")][0])
        print(tr, "ms/step", round(d["ms_per_step"],4), "tok/s", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],4), d.get("dp_check"), d["roofline"]["class_ms_per_step"])
    except Exception as e: print(tr, "failed", e)
PY
