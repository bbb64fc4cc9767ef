#!/bin/bash
# quick GPU check used between changes: GPU tests, Large bench (no CPU baseline), recurrence phase trace, Medium bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print('large', round(d['ms_per_step'],4), round(d['value']), d['roofline']['class_ms_per_step'])"; tail -2 gpurun_out/bench_quick.err
timeout 120 python tools/rec_trace.py large > gpurun_out/rec_trace_large.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/rec_trace_large.json'))
for k in d: print(k, round(d[k]['clk_per_step']), {a: round(b) for a, b in d[k]['phase_offsets_clk'].items()})"
timeout 300 python bench.py --config medium --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_medium.json 2> gpurun_out/bench_medium.err; python -c "
import json; d=json.load(open('gpurun_out/bench_medium.json')); print('medium', round(d['ms_per_step'],4), round(d['value']))"
