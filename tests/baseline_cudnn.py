#!/usr/bin/env python
"""Measure the bar BASELINE.json's north_star names: the reference's `--lstm_type pytorch`
train step on the SAME B200 (cuDNN nn.LSTM + cuBLAS addmm + eager softmax, torch defaults:
cudnn.allow_tf32=True, matmul TF32 off), via oracle/torch_port.py (the reference's own torch
calls; the reference scripts themselves cannot travel to the GPU box).  Also records the
reference path's own CPU-fp32 vs GPU discrepancy, which is what "a stated fp32 tolerance"
is anchored to.  Test infrastructure: writes JSON to stdout, never imported by the product.

    python tests/baseline_cudnn.py [--configs small,medium,large] [--steps 30]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import CONFIGS
from oracle import lstm_lm_oracle as O
from oracle import torch_port as P


def time_gpu(c, steps, warmup, tf32):
    torch.backends.cudnn.allow_tf32 = tf32
    model = P.TorchLstmLm(c["V"], c["H"], c["L"], c["p"], c["winit"], seed=1).cuda()
    model.train()
    data = P.synthetic_batches(c["V"], c["B"], c["T"], steps + warmup)
    states = model.zero_state(c["B"])
    for i, (x, y) in enumerate(data):
        if i == warmup:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
        # the reference hands CPU tensors to the model; W[x] does the H2D implicitly (main.py:111)
        _, _, states = P.train_step(model, x.cuda(), y.cuda(), states, c["lr"], c["clip"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    wall = (time.perf_counter() - t0) / steps * 1e3
    return {"ms_per_step_events": ms, "ms_per_step_wall": wall, "tokens_per_s": c["T"] * c["B"] / (wall * 1e-3),
            "cudnn_allow_tf32": tf32}


def discrepancy(c):
    """eval-mode logits of the reference path: CPU fp32 vs GPU (default flags) vs fp64 oracle."""
    torch.backends.cudnn.allow_tf32 = True
    m = P.TorchLstmLm(c["V"], c["H"], c["L"], c["p"], c["winit"], seed=1)
    m.eval()
    x, y = P.synthetic_batches(c["V"], c["B"], c["T"], 1)[0]
    with torch.no_grad():
        cpu, _ = m(x, m.zero_state(c["B"]))
        mg = m.cuda()
        gpu, _ = mg(x.cuda(), mg.zero_state(c["B"]))
        torch.backends.cudnn.allow_tf32 = False
        gpu_fp32, _ = mg(x.cuda(), mg.zero_state(c["B"]))
    params = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in m.reference_state_dict().items()}
    sc, _, _ = O.model_fwd(params, x.numpy(), O.zero_states(c["L"], c["B"], c["H"], np.float64), c["L"])
    scale = float(np.abs(sc).max())
    d = lambda a: float(np.abs(a.cpu().numpy().astype(np.float64) - sc).max())
    return {"logit_scale": scale, "cpu_fp32_vs_fp64": d(cpu), "gpu_default_tf32_vs_fp64": d(gpu),
            "gpu_tf32_off_vs_fp64": d(gpu_fp32)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="small,medium,large")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=10)
    a = ap.parse_args()
    out = {"torch": torch.__version__, "cudnn": torch.backends.cudnn.version(), "gpu": torch.cuda.get_device_name(0),
           "host_cpus": os.cpu_count()}
    for name in a.configs.split(","):
        c = CONFIGS[name]
        out[name] = {"tf32_default": time_gpu(c, a.steps, a.warmup, True), "tf32_off": time_gpu(c, a.steps, a.warmup, False),
                     "discrepancy": discrepancy(c)}
    print(json.dumps(out, indent=1))
