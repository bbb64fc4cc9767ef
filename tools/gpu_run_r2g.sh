#!/bin/bash
mkdir -p gpurun_out
export ZRB_ERROR_REPORT=gpurun_out/r02_error_at_baseline_configs.json ZRB_ERROR_REPORT2=gpurun_out/r02_error_fixture_cases.json
ZRB_TEST_ENGINES=tc timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -x 2>&1 | tail -15
for mode in split nosplit; do
  if [ $mode = nosplit ]; then export ZRB_REC_NOSPLIT=1; else unset ZRB_REC_NOSPLIT; fi
  timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$mode.json')); print('$mode lazy', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['roofline']['class_ms_per_step'])" || tail -3 gpurun_out/bench_$mode.err
  timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gpu-baseline --strict-update > gpurun_out/bench_${mode}_strict.json 2> gpurun_out/bench_${mode}_strict.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${mode}_strict.json')); print('$mode strict', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))" || tail -3 gpurun_out/bench_${mode}_strict.err
  timeout 120 python tools/rec_trace.py large > gpurun_out/rec_trace_$mode.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/rec_trace_$mode.json')); print('$mode trace', {k: round(v['clk_per_step']) for k,v in d.items()}); print({k: round(x) for k,x in d['fwd']['phase_offsets_clk'].items()}); print({k: round(x) for k,x in d['bwd']['phase_offsets_clk'].items()})"
done
