#!/usr/bin/env python
"""Throughput of the reference's OWN training loop (main.py:107-117 verbatim: zero_grad, detach, forward,
nll_loss, backward, clip_grad_norm_, per-parameter SGD) when `model.Model` is the zaremba_b200 drop-in, next to
the same loop on the reference's cuDNN path (oracle/torch_port.py).  CPU [T,B] views in, like main.py."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn
import zaremba_b200
from oracle import torch_port as P
from bench import CONFIGS

name = sys.argv[1] if len(sys.argv) > 1 else "large"
steps, warm = 40, 8
c = CONFIGS[name]
V, H, L, T, B = c["V"], c["H"], c["L"], c["T"], c["B"]
data = P.synthetic_batches(V, B, T, steps + warm)


def nll_loss(scores, y):                       # main.py:77-84
    batch_size = y.size(1)
    expscores = scores.exp()
    probabilities = expscores / expscores.sum(1, keepdim=True)
    answerprobs = probabilities[range(len(y.reshape(-1))), y.reshape(-1)]
    return torch.mean(-torch.log(answerprobs) * batch_size)


def loop(model, state_init, detach):
    states = state_init(B)
    model.train()
    t0 = None
    for i, (x, y) in enumerate(data):
        if i == warm:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        model.zero_grad()
        states = detach(states)
        scores, states = model(x, states)
        loss = nll_loss(scores, y)
        loss.backward()
        with torch.no_grad():
            nn.utils.clip_grad_norm_(model.parameters(), c["clip"])
            for param in model.parameters():
                param -= c["lr"] * param.grad
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps


torch.manual_seed(1)
ours = zaremba_b200.Model(V, H, L, c["p"], c["winit"]).to("cuda")
dt_ours = loop(ours, ours.state_init, ours.detach)
ref = P.TorchLstmLm(V, H, L, c["p"], c["winit"], seed=1).cuda()
ref_model = lambda x, st: ref(x.cuda(), st)
class _Wrap(nn.Module):
    def __init__(s): super().__init__(); s.m = ref
    def forward(s, x, st): return s.m(x.cuda(), st)
w = _Wrap()
dt_ref = loop(w, ref.zero_state, lambda st: [(h.detach(), cc.detach()) for h, cc in st])
print(json.dumps({"config": name, "loop": "main.py:107-117 verbatim", "dropin_ms_per_step": dt_ours * 1e3,
                  "dropin_tokens_per_s": T * B / dt_ours, "reference_cudnn_ms_per_step": dt_ref * 1e3,
                  "reference_cudnn_tokens_per_s": T * B / dt_ref, "speedup": dt_ref / dt_ours}))
