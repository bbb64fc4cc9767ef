#!/bin/bash
# watchdog fault-injection tests + a perf A/B line (the watchdog must cost nothing on the fast path)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_watchdog.py tests/test_gpu_parity.py -m gpu -q -x -k "watchdog or lost_arrival or healthy or launch_modes or lazy or unit_abi" 2>&1 | tail -15
for i in 1 2; do
timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_wd_$i.json 2> gpurun_out/bench_wd_$i.err
python -c "
import json; d=json.load(open('gpurun_out/bench_wd_$i.json')); print('large', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))" || tail -3 gpurun_out/bench_wd_$i.err
done
timeout 200 python bench.py --config small --steps 300 --warmup 20 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_wd_small.json 2> gpurun_out/bench_wd_small.err
python -c "
import json; d=json.load(open('gpurun_out/bench_wd_small.json')); print('small', round(d['ms_per_step'],4))"
