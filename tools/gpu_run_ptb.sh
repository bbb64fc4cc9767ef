#!/bin/bash
# The README's three single-model recipes (README.md:20-27) on the real PTB id fixture: ours (fused Trainer) and the
# reference's --lstm_type pytorch path on the same GPU (cuDNN, torch port), same seed.  Results -> gpurun_out/ptb_*.json
mkdir -p gpurun_out
run() { # recipe impl timeout
  timeout $3 python tools/train_ptb.py --recipe $1 --impl $2 --json gpurun_out/ptb_$1_$2.json > gpurun_out/ptb_$1_$2.log 2>&1
  echo "== $1 $2 rc=$? $(grep -E 'Test set' gpurun_out/ptb_$1_$2.log) $(grep -E 'Epoch' gpurun_out/ptb_$1_$2.log | tail -1)"
}
run small ours 300
run small cudnn 400
run medium ours 400
run large ours 600
run medium cudnn 700
run large cudnn 900
