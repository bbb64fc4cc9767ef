#!/bin/bash
mkdir -p gpurun_out
for mode in overlap nooverlap; do
  if [ $mode = nooverlap ]; then export ZRB_NO_OVERLAP=1; else unset ZRB_NO_OVERLAP; fi
  timeout 120 python tools/rec_trace.py large > gpurun_out/rec_trace_$mode.json 2> gpurun_out/rec_trace_$mode.err; python -c "
import json; d=json.load(open('gpurun_out/rec_trace_$mode.json')); print('$mode', {k: round(v['clk_per_step']) for k,v in d.items()}, {k: round(x) for k,x in d['bwd']['phase_offsets_clk'].items()})"
  ZRB_PROF_KEEP_PDL=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_keep_$mode.json 2> gpurun_out/bench_keep_$mode.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_keep_$mode.json')); print('$mode', round(d['ms_per_step'],4), d['roofline']['class_ms_per_step'])"
done
