#!/usr/bin/env python
"""Error of the tcgen05 engine vs the fp64 oracle at the Large config (eval-mode logits, train-mode grads),
next to the reference's own cuDNN path (TF32) measured the same way.  Prints JSON."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zaremba_b200
from oracle import lstm_lm_oracle as O
from oracle import torch_port as P
from bench import CONFIGS

c = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "large"]
V, H, L, T, B = c["V"], c["H"], c["L"], c["T"], c["B"]
x, y = P.synthetic_batches(V, B, T, 1)[0]
out = {}
torch.manual_seed(1)
m = zaremba_b200.Model(V, H, L, 0.0, c["winit"]).cuda(); m.train()
params = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in m.named_parameters()}
sc, st, cache = O.model_fwd(params, x.numpy(), O.zero_states(L, B, H, np.float64), L)
grads = O.model_bwd(params, cache, O.nll_loss_bwd(sc, y.numpy()), L)
scores, _ = m(x, m.state_init(B))
e = scores.exp(); p = e / e.sum(1, keepdim=True)
loss = torch.mean(-torch.log(p[torch.arange(T * B, device="cuda"), y.reshape(-1).cuda()]) * B)
loss.backward()
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
out["ours_tc"] = {"logits_rel_to_scale": rel(scores.detach().cpu().numpy(), sc), "loss_rel": abs(loss.item() - O.nll_loss(sc, y.numpy())) / O.nll_loss(sc, y.numpy()),
                  "grads_rel_to_scale": {k: rel(v.grad.cpu().numpy(), grads[k]) for k, v in m.named_parameters()}}
# the reference's own GPU path (torch nn.LSTM on cuDNN, default flags) on the same weights
ref = P.TorchLstmLm(V, H, L, 0.0, c["winit"]).cuda(); ref.train()
ref.load_reference_state_dict({k: v.detach() for k, v in m.named_parameters()})
logits, _ = ref(x.cuda(), ref.zero_state(B))
l2 = P.softmax_nll_times_batch(logits, y.cuda()); l2.backward()
out["reference_cudnn_tf32"] = {"logits_rel_to_scale": rel(logits.detach().cpu().numpy(), sc),
                               "grads_rel_to_scale": {k: rel(v.grad.cpu().numpy(), grads[k]) for k, v in ref.reference_state_dict().items()}}
print(json.dumps(out, indent=1))
