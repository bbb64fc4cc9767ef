#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log; tail -4 gpurun_out/pytest_all.log
timeout 120 python tools/rec_trace.py large > gpurun_out/rec_trace_large.json 2> gpurun_out/rec_trace.err; cat gpurun_out/rec_trace_large.json; tail -2 gpurun_out/rec_trace.err
timeout 300 python bench.py --engine tc --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tc_v4.json 2> gpurun_out/bench_tc_v4.err; cat gpurun_out/bench_tc_v4.json; tail -3 gpurun_out/bench_tc_v4.err
# launch list of one steady-state step (skip construction + warm-up launches)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tc.csv python bench.py --engine tc --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
for k in lstm_rec_fwd_kernel lstm_rec_bwd_kernel gemm_f16_tc_kernel clip_sgd_update_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 2 -f -o gpurun_out/prof_$k python bench.py --engine tc --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
