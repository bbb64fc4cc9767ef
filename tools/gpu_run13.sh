#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_tc_v9.json 2> gpurun_out/bench_tc_v9.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tc_v9.json')); print('large', round(d['ms_per_step'],4), round(d['value']), d['roofline']['class_ms_per_step'])"; tail -2 gpurun_out/bench_tc_v9.err
