#!/bin/bash
mkdir -p gpurun_out
for bn in 256 128; do
  for ab in "0 0" "0 1" "1 1"; do
    ZRB_GEMM_BN=$bn timeout 120 python tools/test_gemm_tc.py $ab > gpurun_out/gemm_bn${bn}_${ab// /}.json 2>/dev/null; echo "bn$bn $ab $(tail -1 gpurun_out/gemm_bn${bn}_${ab// /}.json)"
  done
done
python - <<'PY'
import json
for ab in ("00","01","11"):
    a=json.loads(open(f"gpurun_out/gemm_bn256_{ab}.json").readline()); b=json.loads(open(f"gpurun_out/gemm_bn128_{ab}.json").readline())
    for ca,cb in zip(a["cases"],b["cases"]):
        if ca["shape"][0]>=700: print(ab, ca["shape"], "bn256", ca.get("us"), ca.get("tflops"), "| bn128", cb.get("us"), cb.get("tflops"), cb.get("ok"))
PY
ZRB_GEMM_BN=128 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_bn128.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_bn128.json')); print('bn128 step', round(d['ms_per_step'],4), d['roofline']['class_ms_per_step'])"
