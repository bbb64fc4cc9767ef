"""Torch restatement of the reference's `--lstm_type pytorch` train/eval step.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/lstm_lm_oracle.py for the rules).

The reference is three Python scripts whose arithmetic lives in PyTorch
(`nn.LSTM` -> oneDNN on CPU, cuDNN on GPU; `addmm`; eager softmax).  Those
scripts cannot travel to the GPU box, torch can.  This port issues the SAME
torch library calls in the same order, so timing it on the box's host cores is
the closest stand-in for `python main.py --device cpu` (`cpu_baseline.kind =
"port"`), and timing it on `cuda` gives the cuDNN bar that BASELINE.json's
north_star asks us to beat by 2x.

Restated reference lines (/root/reference):
  model.py:76-92    parameter set, registration order, U(-winit, winit) init,
                    including nn.LSTM's own constructor draws that precede it
  model.py:103-110  forward: W[x] -> dropout -> (nn.LSTM -> dropout) x L -> addmm
  main.py:77-84     naive softmax NLL * batch_size
  main.py:109-117   zero_grad, detach, fwd, loss, bwd, clip_grad_norm_, p -= lr*g

Checked against the golden fixtures by tests/test_torch_port.py.
"""
from __future__ import annotations

import math

import torch
from torch import nn


class TorchLstmLm(nn.Module):
    def __init__(self, vocab, hidden, layers, dropout, winit, seed=None):
        super().__init__()
        if seed is not None:
            torch.manual_seed(seed)
        self.vocab, self.hidden, self.layers, self.p = vocab, hidden, layers, dropout
        self.emb_w = nn.Parameter(torch.empty(vocab, hidden))
        # nn.LSTM's constructor consumes RNG (its own reset_parameters) before the
        # model-wide re-init, exactly as in the reference (model.py:84 then :88).
        self.cells = nn.ModuleList(nn.LSTM(hidden, hidden) for _ in range(layers))
        self.out_w = nn.Parameter(torch.empty(vocab, hidden))
        self.out_b = nn.Parameter(torch.empty(vocab))
        self.drop = nn.Dropout(dropout)
        # model.py:90-92: one uniform_ per parameter in the REFERENCE's registration
        # order (embed, rnns.*, fc) -- not this module's own attribute order.
        for prm in self.reference_state_dict().values():
            nn.init.uniform_(prm, -winit, winit)

    # name map to the reference's state_dict keys
    def reference_state_dict(self):
        out = {"embed.W": self.emb_w}
        for l, cell in enumerate(self.cells):
            for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
                out[f"rnns.{l}.{k}"] = getattr(cell, k)
        out["fc.W"] = self.out_w
        out["fc.b"] = self.out_b
        return out

    def load_reference_state_dict(self, sd):
        with torch.no_grad():
            for k, dst in self.reference_state_dict().items():
                dst.copy_(torch.as_tensor(sd[k]))

    def zero_state(self, batch):
        dev = self.emb_w.device
        return [(torch.zeros(1, batch, self.hidden, device=dev),
                 torch.zeros(1, batch, self.hidden, device=dev)) for _ in self.cells]

    def forward(self, x, states):
        a = self.drop(self.emb_w[x])
        new_states = []
        for cell, st in zip(self.cells, states):
            a, st2 = cell(a, st)
            new_states.append(st2)
            a = self.drop(a)
        logits = torch.addmm(self.out_b, a.view(-1, self.hidden), self.out_w.t())
        return logits, new_states


def softmax_nll_times_batch(logits, y):
    """main.py:77-84 (no max subtraction, like the reference)."""
    e = logits.exp()
    p = e / e.sum(1, keepdim=True)
    tgt = y.reshape(-1)
    picked = p[torch.arange(tgt.numel(), device=p.device), tgt]
    return torch.mean(-torch.log(picked) * y.size(1))


def train_step(model, x, y, states, lr, max_norm):
    """main.py:109-117."""
    model.zero_grad()
    states = [(h.detach(), c.detach()) for h, c in states]
    logits, states = model(x, states)
    loss = softmax_nll_times_batch(logits, y)
    loss.backward()
    with torch.no_grad():
        norm = nn.utils.clip_grad_norm_(model.parameters(), max_norm)
        for prm in model.parameters():
            prm -= lr * prm.grad
    return loss, norm, states


def synthetic_batches(vocab, batch, seq, n_batches, seed=2):
    """SURVEY 8d synthetic PTB-shaped tokens: data[B, T*n+1]; batch i = the
    transposed (non-contiguous) [T,B] window, targets shifted by one."""
    g = torch.Generator().manual_seed(seed)
    data = torch.randint(0, vocab, (batch, seq * n_batches + 1), generator=g, dtype=torch.int64)
    return [(data[:, i * seq:(i + 1) * seq].t(), data[:, i * seq + 1:(i + 1) * seq + 1].t())
            for i in range(n_batches)]


def time_cpu_train_steps(vocab, hidden, layers, batch, seq, dropout, winit, lr, max_norm,
                         steps, warmup, threads=None, device="cpu"):
    """Wall-clock train steps of the port; returns (sec_per_step, tokens_per_sec, threads)."""
    import time
    if threads:
        torch.set_num_threads(threads)
    model = TorchLstmLm(vocab, hidden, layers, dropout, winit, seed=1).to(device)
    model.train()
    data = synthetic_batches(vocab, batch, seq, steps + warmup)
    states = model.zero_state(batch)
    sync = torch.cuda.synchronize if device != "cpu" else (lambda: None)
    t0 = 0.0
    for i, (x, y) in enumerate(data):
        if i == warmup:
            sync()
            t0 = time.perf_counter()
        if device != "cpu":
            x, y = x.to(device), y.to(device)
        _, _, states = train_step(model, x, y, states, lr, max_norm)
    sync()
    dt = (time.perf_counter() - t0) / max(1, steps)
    return dt, batch * seq / dt, torch.get_num_threads()
