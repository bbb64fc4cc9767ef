#!/bin/bash
# round-2 iteration A: full GPU test-suite (incl. the baseline-config tests), bench with/without the PDL overlap,
# profiler environment probe, ncu launch list under smoke()
mkdir -p gpurun_out
export ZRB_ERROR_REPORT=gpurun_out/r02_error_at_baseline_configs.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log; tail -15 gpurun_out/pytest_all.log
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_overlap.json 2> gpurun_out/bench_overlap.err; cut -c1-2500 gpurun_out/bench_overlap.json; tail -3 gpurun_out/bench_overlap.err
ZRB_NO_OVERLAP=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_nooverlap.json 2> gpurun_out/bench_nooverlap.err; python - <<'PY'
import json
for n in ("overlap","nooverlap"):
    try:
        d=json.load(open(f"gpurun_out/bench_{n}.json")); print(n, round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), d["gpu_launches"], d["roofline"]["class_ms_per_step"])
    except Exception as e: print(n, "failed", e)
PY
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 5 python -c "
import os; print({k:v for k,v in os.environ.items() if any(s in k for s in ('INJECT','NSIGHT','NV_','CUDA_'))})" > gpurun_out/ncu_env_probe.txt 2>&1; tail -3 gpurun_out/ncu_env_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_smoke.csv python __graft_entry__.py smoke > gpurun_out/ncu_smoke.log 2>&1; echo "ncu smoke rc=$?"; grep -c "lstm_rec_bwd" gpurun_out/launches_smoke.csv; tail -3 gpurun_out/ncu_smoke.log
