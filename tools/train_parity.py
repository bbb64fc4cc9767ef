#!/usr/bin/env python
"""Training-curve and validation-perplexity parity on the GPU box: the reference's `--lstm_type pytorch`
path (oracle/torch_port.py on cuda: cuDNN nn.LSTM, eager loss, clip, SGD) and zaremba_b200.Trainer, same
initial weights, same learnable synthetic corpus (PTB itself cannot travel to the box), same schedule.
Dropout 0 so both runs are deterministic functions of the data.  Prints JSON with the two loss curves'
divergence and the two validation perplexities (main.py:86-95 semantics).

    python tools/train_parity.py [small|medium] [steps]
"""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zaremba_b200
from oracle import torch_port as P
from bench import CONFIGS

name = sys.argv[1] if len(sys.argv) > 1 else "small"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
c = dict(CONFIGS[name]); c["p"] = 0.0
V, H, L, T, B = c["V"], c["H"], c["L"], c["T"], c["B"]

def corpus(n, seed):
    """Learnable token stream: a sparse random bigram model with a Zipf-like unigram back-off."""
    rng = np.random.default_rng(seed)
    nxt = rng.integers(0, V, size=(V, 4))                      # 4 likely successors per token
    zipf = (rng.zipf(1.3, size=n) - 1) % V
    out = np.empty(n, dtype=np.int64); out[0] = 0
    r = rng.random(n); pick = rng.integers(0, 4, size=n)
    for i in range(1, n):
        out[i] = nxt[out[i - 1], pick[i]] if r[i] < 0.8 else zipf[i]
    return out

trn = zaremba_b200.minibatch(corpus(B * (T * steps + 1) + 7, 1).reshape(-1, 1), B, T)[:steps]
vld = zaremba_b200.minibatch(corpus(B * (T * 20 + 1) + 3, 2).reshape(-1, 1), B, T)

torch.manual_seed(1)
ref = P.TorchLstmLm(V, H, L, 0.0, c["winit"]).cuda(); ref.train()
ours = zaremba_b200.Model(V, H, L, 0.0, c["winit"]).cuda(); ours.train()
ours.load_state_dict({k: v.detach().clone() for k, v in ref.reference_state_dict().items()})
tr = zaremba_b200.Trainer(ours, B, T)

ref32 = P.TorchLstmLm(V, H, L, 0.0, c["winit"]).cuda(); ref32.train()
ref32.load_reference_state_dict({k: v.detach().clone() for k, v in ref.reference_state_dict().items()})
ref_losses, our_losses, ref32_losses = [], [], []
states = ref.zero_state(B)
states32 = ref32.zero_state(B)
lr = c["lr"]
for i, (x, y) in enumerate(trn):
    torch.backends.cudnn.allow_tf32 = True          # torch default: what `main.py --device gpu` runs
    l_ref, _, states = P.train_step(ref, x.cuda(), y.cuda(), states, lr, c["clip"])
    torch.backends.cudnn.allow_tf32 = False         # the reference's own fp32 cuDNN path, for scale
    l_r32, _, states32 = P.train_step(ref32, x.cuda(), y.cuda(), states32, lr, c["clip"])
    l_our, _ = tr.train_step(x.contiguous().cuda(), y.contiguous().cuda(), lr, c["clip"])
    ref_losses.append(l_ref.item() / B); our_losses.append(l_our.item() / B); ref32_losses.append(l_r32.item() / B)
torch.backends.cudnn.allow_tf32 = True

def ref_ppl(ref=ref):
    ref.eval(); st = ref.zero_state(B); ls = []
    with torch.no_grad():
        for x, y in vld:
            logits, st = ref(x.cuda(), st)
            ls.append(P.softmax_nll_times_batch(logits, y.cuda()).item() / B)
    return math.exp(float(np.mean(ls)))

ours.eval()
out = {"config": name, "steps": steps, "tokens": steps * T * B,
       "train_loss_first": [ref_losses[0], our_losses[0]], "train_loss_last10_mean": [float(np.mean(ref_losses[-10:])), float(np.mean(our_losses[-10:]))],
       "max_abs_loss_gap": float(np.max(np.abs(np.array(ref_losses) - np.array(our_losses)))),
       "max_rel_loss_gap": float(np.max(np.abs(np.array(ref_losses) - np.array(our_losses)) / np.array(ref_losses))),
       "valid_ppl_reference_cudnn": ref_ppl(), "valid_ppl_ours": tr.perplexity(vld),
       "valid_ppl_reference_cudnn_tf32_off": ref_ppl(ref32),
       "reference_tf32_vs_fp32_max_rel_loss_gap": float(np.max(np.abs(np.array(ref_losses) - np.array(ref32_losses)) / np.array(ref_losses))),
       "loss_curve_every_25": [[round(a, 4), round(b, 4)] for a, b in list(zip(ref_losses, our_losses))[::25]]}
out["reference_tf32_vs_fp32_valid_ppl_rel_gap"] = abs(out["valid_ppl_reference_cudnn_tf32_off"] - out["valid_ppl_reference_cudnn"]) / out["valid_ppl_reference_cudnn"]
out["valid_ppl_rel_gap"] = abs(out["valid_ppl_ours"] - out["valid_ppl_reference_cudnn"]) / out["valid_ppl_reference_cudnn"]
print(json.dumps(out, indent=1))
