#!/bin/bash
mkdir -p gpurun_out
export ZRB_ERROR_REPORT=gpurun_out/r02_error_at_baseline_configs.json
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log; tail -12 gpurun_out/pytest_all.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_overlap.json 2> gpurun_out/bench_overlap.err; tail -3 gpurun_out/bench_overlap.err
ZRB_NO_OVERLAP=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_nooverlap.json 2> gpurun_out/bench_nooverlap.err
python - <<'PY'
import json
for n in ("overlap","nooverlap"):
    try:
        d=json.load(open(f"gpurun_out/bench_{n}.json")); print(n, round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), d["gpu_launches"], d.get("vs_baseline"), d["roofline"]["class_ms_per_step"])
    except Exception as e: print(n, "failed", e)
PY
