#!/bin/bash
# same-box A/B of the working tree against a prebuilt copy of an earlier commit under _ab_old/ (made with `git archive`),
# after the parity + watchdog tests of the working tree
mkdir -p gpurun_out
ZRB_TEST_ENGINES=tc timeout 400 python -m pytest tests/test_gpu_watchdog.py tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q 2>&1 | tail -5
for i in 1 2; do
  for w in new old; do
    if [ $w = old ]; then cd _ab_old; fi
    timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-gpu-baseline > /tmp/ab_$w.json 2> /tmp/ab_$w.err
    python -c "
import json; d=json.load(open('/tmp/ab_$w.json')); print('$w large', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['roofline']['class_ms_per_step']['rec_fwd'], d['roofline']['class_ms_per_step']['rec_bwd'])" || tail -3 /tmp/ab_$w.err
    if [ $w = old ]; then cd ..; fi
  done
done
