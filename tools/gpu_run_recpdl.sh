#!/bin/bash
# recurrence kernels as programmatic dependents of the GEMM before them: ZRB_REC_PDL = 1 (with the cooperative attribute),
# 2 (plain checked cluster launch), 0 (off); "plain" = ZRB_NO_COOP=1 without the programmatic attribute
mkdir -p gpurun_out
for mode in ${MODES:-plain 2 0 plain 2}; do
  if [ $mode = plain ]; then export ZRB_NO_COOP=1 ZRB_REC_PDL=0; else unset ZRB_NO_COOP; export ZRB_REC_PDL=$mode; fi
  ZRB_VERBOSE=1 timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_recpdl_$mode.json 2> gpurun_out/bench_recpdl_$mode.err
  grep -h "zrb\]" gpurun_out/bench_recpdl_$mode.err | sort | uniq -c
  python -c "
import json; d=json.load(open('gpurun_out/bench_recpdl_$mode.json')); print('mode $mode', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))" || tail -3 gpurun_out/bench_recpdl_$mode.err
done
