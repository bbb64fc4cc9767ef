#!/bin/bash
# round-2 profiling at HEAD: ncu launch list of steady-state steps + full captures of the dominant kernels.
# (ncu forces ZRB_NO_COOP behaviour automatically: the library detects the profiler and drops the cooperative attribute.)
mkdir -p gpurun_out; rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches_*.csv
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --strict-update"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_tc.csv $B > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"; grep -c "zrb::" gpurun_out/launches_tc.csv
for k in lstm_rec_fwd_kernel lstm_rec_bwd_kernel gemm_f16_tc_kernel update_pack softmax_nll; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 4 -f -o gpurun_out/prof_$k $B > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep | wc -l
