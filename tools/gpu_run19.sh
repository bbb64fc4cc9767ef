#!/bin/bash
mkdir -p gpurun_out
export ZRB_GEMM_MT=2
for ab in "0 0" "1 1"; do
  set -- $ab
  timeout 120 python tools/test_gemm_tc.py $1 $2 > gpurun_out/gemm_mt2_$1$2.json 2> gpurun_out/gemm_mt2_$1$2.err; echo "gemm $1 $2 rc=$? $(tail -1 gpurun_out/gemm_mt2_$1$2.json | cut -c1-60)"
  python - gpurun_out/gemm_mt2_$1$2.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line)
        print([(c['shape'],c.get('us'),c.get('tflops'),c.get('ok')) for c in d['cases'] if (c['shape'][0]>=700 or c['shape'][1]>=1500) or not c.get('ok')])
PY
done
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_mt2.json 2> gpurun_out/bench_mt2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_mt2.json')); print('large mt2', round(d['ms_per_step'],4), round(d['value']), d['roofline']['class_ms_per_step'])"
