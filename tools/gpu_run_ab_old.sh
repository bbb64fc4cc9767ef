#!/bin/bash
# same-box A/B of the working tree against a prebuilt copy of HEAD under _ab_old/ (made with `git archive`)
mkdir -p gpurun_out
for i in 1 2; do
  for w in new old; do
    if [ $w = old ]; then cd _ab_old; fi
    timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-gpu-baseline > /tmp/ab_$w.json 2> /tmp/ab_$w.err
    python -c "
import json; d=json.load(open('/tmp/ab_$w.json')); print('$w large', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['roofline']['class_ms_per_step']['rec_fwd'], d['roofline']['class_ms_per_step']['rec_bwd'])" || tail -3 /tmp/ab_$w.err
    if [ $i = 1 ]; then
    timeout 200 python bench.py --config small --steps 300 --warmup 20 --no-cpu-baseline --no-gpu-baseline > /tmp/ab_s_$w.json 2> /tmp/ab_s_$w.err
    python -c "
import json; d=json.load(open('/tmp/ab_s_$w.json')); print('$w small', round(d['ms_per_step'],4), d['roofline']['class_ms_per_step']['rec_fwd'], d['roofline']['class_ms_per_step']['rec_bwd'])" || tail -3 /tmp/ab_s_$w.err
    fi
    if [ $w = old ]; then cd ..; fi
  done
done
