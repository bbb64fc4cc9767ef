#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_tc_v10.json 2> gpurun_out/bench_tc_v10.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tc_v10.json')); print('large', round(d['ms_per_step'],4), round(d['value']), d['roofline']['class_ms_per_step'])"; tail -2 gpurun_out/bench_tc_v10.err
timeout 300 python tools/bench_dropin.py large > gpurun_out/dropin_large.json 2> gpurun_out/dropin_large.err; cat gpurun_out/dropin_large.json; tail -3 gpurun_out/dropin_large.err
timeout 300 python tools/bench_dropin.py medium > gpurun_out/dropin_medium.json 2> gpurun_out/dropin_medium.err; cat gpurun_out/dropin_medium.json
