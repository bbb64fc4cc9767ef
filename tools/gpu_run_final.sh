#!/bin/bash
# round-end style validation: full GPU tests (both engines), smoke, default bench, ncu launch list + captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log; tail -3 gpurun_out/pytest_all.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cat gpurun_out/bench_final.json | cut -c1-3000; tail -2 gpurun_out/bench_final.err
timeout 300 python bench.py --config medium --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_medium.json 2> gpurun_out/bench_medium.err; python -c "
import json; d=json.load(open('gpurun_out/bench_medium.json')); print('medium', round(d['ms_per_step'],4), round(d['value']), d['roofline']['class_ms_per_step'])"
# launch list of steady-state steps; the 4-CTA-cluster cooperative kernel cannot be launched under ncu -> profile everything else
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(?!lstm_rec_bwd).*" -c 400 --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"; grep -c "zrb::" gpurun_out/launches_tc.csv
for k in lstm_rec_fwd_kernel gemm_f16_tc_kernel softmax_nll_kernel update_pack; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 3 -f -o gpurun_out/prof_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ZRB_NO_COOP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_bwd_kernel -s 2 -c 2 -f -o gpurun_out/prof_lstm_rec_bwd_kernel python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_lstm_rec_bwd_kernel.log 2>&1; echo "ncu rec_bwd rc=$?"
timeout 120 python tools/rec_trace.py large > gpurun_out/rec_trace_large.json 2>/dev/null
ls -la gpurun_out/*.ncu-rep | wc -l
