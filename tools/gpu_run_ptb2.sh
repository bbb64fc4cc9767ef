#!/bin/bash
# Large recipe again at HEAD (K-split kernels, lazy update off as in the Trainer default) + the reference's cuDNN path
# with other seeds (seed 0 diverged): stop at the first seed that trains
mkdir -p gpurun_out
timeout 500 python tools/train_ptb.py --recipe large --impl ours --json gpurun_out/ptb_large_ours_head.json > gpurun_out/ptb_large_ours_head.log 2>&1
echo "== large ours (HEAD) rc=$? $(grep -E 'Test set' gpurun_out/ptb_large_ours_head.log) $(grep -E 'Epoch' gpurun_out/ptb_large_ours_head.log | tail -1)"
for seed in 1 2 3; do
  timeout 700 python tools/train_ptb.py --recipe large --impl cudnn --seed $seed --json gpurun_out/ptb_large_cudnn_seed$seed.json > gpurun_out/ptb_large_cudnn_seed$seed.log 2>&1
  rc=$?; echo "== large cudnn seed $seed rc=$rc $(grep -E 'NON-FINITE|Test set' gpurun_out/ptb_large_cudnn_seed$seed.log) $(grep -E 'Epoch' gpurun_out/ptb_large_cudnn_seed$seed.log | tail -1)"
  if [ $rc = 0 ]; then break; fi
done
