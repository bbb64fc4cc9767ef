#!/bin/bash
mkdir -p gpurun_out
for mode in strict lazy; do
  flag=""; [ $mode = lazy ] && flag="--lazy_update"
  timeout 300 python tools/train_ptb.py --recipe medium $flag --json gpurun_out/ptb_medium_$mode.json > gpurun_out/ptb_medium_$mode.log 2>&1
  echo "medium $mode rc=$? $(grep 'Test set' gpurun_out/ptb_medium_$mode.log) $(grep Epoch gpurun_out/ptb_medium_$mode.log | tail -1)"
done
python -c "
import json
a=json.load(open('gpurun_out/ptb_medium_strict.json')); b=json.load(open('gpurun_out/ptb_medium_lazy.json'))
print('strict', a['test_ppl'], a['valid_ppl_per_epoch'][:3], round(a['train_tokens_per_s_median_epoch']))
print('lazy  ', b['test_ppl'], b['valid_ppl_per_epoch'][:3], round(b['train_tokens_per_s_median_epoch']))"
