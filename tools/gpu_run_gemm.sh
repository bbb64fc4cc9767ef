#!/bin/bash
# GEMM unit check (all operand-major combinations, ragged shapes) with and without the multicast pairs, then A/B bench
mkdir -p gpurun_out
for mode in mc nomc; do
  if [ $mode = nomc ]; then export ZRB_GEMM_NOMC=1; else unset ZRB_GEMM_NOMC; fi
  for ab in "0 0" "0 1" "1 1" "1 0"; do
    timeout 120 python tools/test_gemm_tc.py $ab > gpurun_out/gemm_${mode}_${ab// /}.json 2> gpurun_out/gemm_${mode}_${ab// /}.err; echo "$mode $ab rc=$? $(tail -1 gpurun_out/gemm_${mode}_${ab// /}.json)"
  done
done
unset ZRB_GEMM_NOMC
python - <<'PY'
import json
for ab in ("00","01","11","10"):
    try:
        a=json.loads(open(f"gpurun_out/gemm_mc_{ab}.json").readline()); b=json.loads(open(f"gpurun_out/gemm_nomc_{ab}.json").readline())
        for ca,cb in zip(a["cases"],b["cases"]):
            if ca["shape"][0]>=128: print(ab, ca["shape"], "mc", ca.get("us"), ca.get("tflops"), ca.get("ok"), "| nomc", cb.get("us"), cb.get("tflops"))
    except Exception as e: print(ab, "failed", e)
PY
bash tools/gpu_run_ab.sh ZRB_GEMM_NOMC
