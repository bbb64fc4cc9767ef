#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/micro/pdl_overlap_bench > gpurun_out/pdl_overlap_micro.txt 2>&1; cat gpurun_out/pdl_overlap_micro.txt
for mode in overlap nooverlap; do
  if [ $mode = nooverlap ]; then export ZRB_NO_OVERLAP=1; else unset ZRB_NO_OVERLAP; fi
  ZRB_PROF_KEEP_PDL=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_keep_$mode.json 2> gpurun_out/bench_keep_$mode.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_keep_$mode.json')); print('$mode', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['roofline']['class_ms_per_step'])"
done
unset ZRB_NO_OVERLAP
ZRB_BWD_PUSH=1 ZRB_TEST_ENGINES=tc timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -x 2>&1 | tail -4
ZRB_BWD_PUSH=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_push.json 2> gpurun_out/bench_push.err
python -c "
import json; d=json.load(open('gpurun_out/bench_push.json')); print('push', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['roofline']['class_ms_per_step'])"
ZRB_BWD_PUSH=1 timeout 120 python tools/rec_trace.py large > gpurun_out/rec_trace_push.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/rec_trace_push.json')); print('push', {k: round(v['clk_per_step']) for k,v in d.items()}, {k: round(x) for k,x in d['bwd']['phase_offsets_clk'].items()})"
