#!/bin/bash
mkdir -p gpurun_out
for cfg in medium small; do
  timeout 300 python bench.py --config $cfg --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$cfg.json')); print('$cfg', round(d['ms_per_step'],4), 'ms', round(d['value']), 'tok/s  e2e', round(d['e2e']['value']), d['roofline']['class_ms_per_step'])"; tail -2 gpurun_out/bench_$cfg.err
done
timeout 600 python tools/train_parity.py small 300 > gpurun_out/train_parity_small.json 2> gpurun_out/train_parity_small.err; head -c 1500 gpurun_out/train_parity_small.json; tail -3 gpurun_out/train_parity_small.err
timeout 600 python tools/train_parity.py medium 150 > gpurun_out/train_parity_medium.json 2> gpurun_out/train_parity_medium.err; head -c 900 gpurun_out/train_parity_medium.json; tail -3 gpurun_out/train_parity_medium.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tc_v8.json 2> gpurun_out/bench_tc_v8.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tc_v8.json')); print('large', d['ms_per_step'], d['value'], d['roofline']['class_ms_per_step'])"
