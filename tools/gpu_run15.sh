#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/micro/tmem_ld_bench > gpurun_out/tmem_ld_bench.csv 2>&1; echo "tmem rc=$?"; grep -v "^#" gpurun_out/tmem_ld_bench.csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_tc_v11.json 2> gpurun_out/bench_tc_v11.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tc_v11.json')); print('large', round(d['ms_per_step'],4), round(d['value']), d['roofline']['class_ms_per_step'])"; tail -2 gpurun_out/bench_tc_v11.err
timeout 120 python tools/rec_trace.py large > gpurun_out/rec_trace_large.json 2>/dev/null; cat gpurun_out/rec_trace_large.json
