#!/bin/bash
# round-end style validation on ONE B200: full GPU test-suite (both engines), smoke under ncu (launch list the driver
# would see), the default bench line (CPU + cuDNN baselines), the strict-update and no-K-split A/B lines, phase traces
mkdir -p gpurun_out
export ZRB_ERROR_REPORT=gpurun_out/r02_error_at_baseline_configs.json ZRB_ERROR_REPORT2=gpurun_out/r02_error_fixture_cases.json
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log; tail -4 gpurun_out/pytest_all.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_smoke.csv python __graft_entry__.py smoke > gpurun_out/ncu_smoke.log 2>&1; echo "ncu smoke rc=$?"; grep -c "lstm_rec_bwd" gpurun_out/launches_smoke.csv; tail -2 gpurun_out/ncu_smoke.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-1200 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gpu-baseline --strict-update > gpurun_out/bench_strict.json 2> gpurun_out/bench_strict.err
ZRB_NO_OVERLAP=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gpu-baseline --strict-update > gpurun_out/bench_strict_nooverlap.json 2> gpurun_out/bench_strict_nooverlap.err

timeout 120 python tools/rec_trace.py large > gpurun_out/rec_trace_final.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 200 python bench.py --config small --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
python - <<'PY'
import json
for n in ("final","strict","strict_nooverlap","reference","small"):
    try:
        d=json.load(open(f"gpurun_out/bench_{n}.json")); print(n, round(d["ms_per_step"],4), round(d["value"]), d.get("vs_baseline"), (d.get("e2e") or {}).get("ms_per_step"), (d.get("cpu_baseline") or {}).get("cores"))
    except Exception as e: print(n, "failed", e)
PY
python -c "
import json; d=json.load(open('gpurun_out/rec_trace_final.json')); print('trace', {k: round(v['clk_per_step']) for k,v in d.items()}); print({k: round(x) for k,x in d['fwd']['phase_offsets_clk'].items()}); print({k: round(x) for k,x in d['bwd']['phase_offsets_clk'].items()})"
