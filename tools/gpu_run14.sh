#!/bin/bash
mkdir -p gpurun_out
ZRB_NO_COOP=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_bwd_kernel -s 2 -c 2 -f -o gpurun_out/prof_lstm_rec_bwd_kernel python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_lstm_rec_bwd_kernel.log 2>&1; echo "ncu bwd rc=$?"; tail -4 gpurun_out/ncu_lstm_rec_bwd_kernel.log
ZRB_NO_COOP=1 timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('nocoop', d['ms_per_step'], d['roofline']['class_ms_per_step']['rec_bwd'])"
