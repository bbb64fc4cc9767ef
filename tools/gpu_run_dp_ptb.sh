#!/bin/bash
# Data-parallel training on real PTB (2 GPUs x batch 20) against ONE process at --batch_size 40 (the same global batch:
# SURVEY 8e says they are the same computation; Small recipe has no dropout, so the two curves must agree to rounding)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
timeout 300 $T tools/train_ptb.py --recipe small --json gpurun_out/ptb_small_dp2.json > gpurun_out/ptb_small_dp2.log 2>&1; echo "small dp2 rc=$? $(grep -E 'Test set' gpurun_out/ptb_small_dp2.log) $(grep Epoch gpurun_out/ptb_small_dp2.log | tail -1)"
timeout 400 python tools/train_ptb.py --recipe small --batch_size 40 --eval_batch_size 20 --json gpurun_out/ptb_small_b40.json > gpurun_out/ptb_small_b40.log 2>&1; echo "small b40 rc=$? $(grep -E 'Test set' gpurun_out/ptb_small_b40.log) $(grep Epoch gpurun_out/ptb_small_b40.log | tail -1)"
timeout 300 $T tools/train_ptb.py --recipe medium --json gpurun_out/ptb_medium_dp2.json > gpurun_out/ptb_medium_dp2.log 2>&1; echo "medium dp2 rc=$? $(grep -E 'Test set' gpurun_out/ptb_medium_dp2.log) $(grep Epoch gpurun_out/ptb_medium_dp2.log | tail -1)"
python - <<'PY'
import json
a=json.load(open("gpurun_out/ptb_small_dp2.json")); b=json.load(open("gpurun_out/ptb_small_b40.json"))
print("valid ppl per epoch dp2:", a["valid_ppl_per_epoch"]); print("valid ppl per epoch b40:", b["valid_ppl_per_epoch"])
print("test", a["test_ppl"], b["test_ppl"], "tok/s", round(a["train_tokens_per_s_median_epoch"]), round(b["train_tokens_per_s_median_epoch"]))
PY
