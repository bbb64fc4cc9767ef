#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log; tail -4 gpurun_out/pytest_all.log
timeout 300 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cat gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
# launch list: the backward cluster kernel is launched cooperatively; ncu serialises kernels, so also try without the cooperative attribute
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 300 --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"; grep -c lstm_rec gpurun_out/launches_tc.csv; tail -3 gpurun_out/ncu_list.log
for k in lstm_rec_fwd_kernel lstm_rec_bwd_kernel gemm_f16_tc_kernel update_pack; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 2 -f -o gpurun_out/prof_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 200 python tools/measure_error.py large > gpurun_out/measure_error_large.json 2> gpurun_out/measure_error.err; cat gpurun_out/measure_error_large.json | head -60
ls -la gpurun_out/*.ncu-rep
