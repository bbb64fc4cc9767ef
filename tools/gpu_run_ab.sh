#!/bin/bash
# A/B of one switch on ONE GPU: tests with the default, then bench default vs `$1=${2:-1}`
mkdir -p gpurun_out
ZRB_TEST_ENGINES=tc timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -x 2>&1 | tail -6
for mode in on off; do
  if [ $mode = off ]; then export $1=${2:-1}; else unset $1; fi
  timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_ab_$mode.json 2> gpurun_out/bench_ab_$mode.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab_$mode.json')); print('$1 feature $mode', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['roofline']['class_ms_per_step'])" || tail -3 gpurun_out/bench_ab_$mode.err
done
