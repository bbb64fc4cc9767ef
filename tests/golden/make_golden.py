#!/usr/bin/env python
"""Mint golden fixtures from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference

The reference has no tests or golden vectors, so parity is
pinned on outputs of the reference itself: this script imports
`/root/reference/model.py` (class `Model`, both `lstm_type`s) and pulls the
functions `nll_loss`, `minibatch`, `perplexity` (main.py:61-95) and
`ensemble_nll_loss` (ensemble.py:97-109) out of the reference files with `ast`
(main.py / ensemble.py run a training job on import, so they cannot be
imported), executes them on seeded inputs with torch on CPU in fp32 and stores
inputs + outputs as compressed .npz files next to this script.  Nothing from
the reference is copied into the repository; the fixtures are data.

`/root/reference` does not exist on the GPU box: tests read only the .npz
files.
"""
import ast
import os
import sys

import numpy as np
import torch

REF = os.environ.get("ZAREMBA_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
from model import Model  # noqa: E402  (the reference's model.py)


def _pull_functions(path, names, extra_globals):
    """exec selected top-level function definitions of a reference script."""
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    mod = ast.Module(body=keep, type_ignores=[])
    ns = {"np": np, "torch": torch, "nn": torch.nn}
    ns.update(extra_globals)
    exec(compile(mod, path, "exec"), ns)
    return ns


class _Args:
    batch_size = 0


_args = _Args()
MAIN = _pull_functions(os.path.join(REF, "main.py"),
                       {"nll_loss", "minibatch", "perplexity"}, {"args": _args})
ENS = _pull_functions(os.path.join(REF, "ensemble.py"), {"ensemble_nll_loss"}, {"args": _args})
ref_nll_loss = MAIN["nll_loss"]
ref_minibatch = MAIN["minibatch"]
ref_perplexity = MAIN["perplexity"]
ref_ensemble_nll_loss = ENS["ensemble_nll_loss"]


def _np(t):
    return t.detach().cpu().numpy().copy()


def _states_in(model, B, H, L, lstm_type, gen, zero):
    states = model.state_init(B)
    if zero:
        return states
    out = []
    for (h, c) in states:
        out.append((torch.empty_like(h).uniform_(-0.5, 0.5, generator=gen),
                    torch.empty_like(c).uniform_(-1.0, 1.0, generator=gen)))
    return out


def step_case(name, V, H, L, T, B, lstm_type, dropout, winit, seed, lr, max_norm,
              steps=1, zero_state=False, store_full=True):
    """`steps` iterations of main.py:109-117 on the reference Model, recording
    everything the oracle / CUDA path must reproduce."""
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1000)
    model = Model(V, H, L, dropout, winit, lstm_type)
    model.train() if dropout > 0 else model.eval()
    out = {"meta": np.array([V, H, L, T, B, steps], dtype=np.int64),
           "lstm_type": np.array(lstm_type), "dropout": np.array(dropout, dtype=np.float64),
           "lr": np.array(lr, dtype=np.float64), "max_norm": np.array(max_norm, dtype=np.float64),
           "winit": np.array(winit, dtype=np.float64), "seed": np.array(seed)}
    if store_full:
        for k, v in model.state_dict().items():
            out["param0/" + k] = _np(v)
    else:
        for k, v in model.state_dict().items():
            a = _np(v).astype(np.float64)
            out["param0_sum/" + k] = np.array([a.sum(), np.abs(a).sum(), (a * a).sum()])
    states = _states_in(model, B, H, L, lstm_type, gen, zero_state)
    for li, (h, c) in enumerate(states):
        out[f"h0/{li}"] = _np(h).reshape(B, H)
        out[f"c0/{li}"] = _np(c).reshape(B, H)
    # the three dropout sites are one nn.Dropout module called L+1 times per forward
    masks = []
    if dropout > 0:
        def hook(mod, inp, res):
            masks.append(_np(res != 0))
        model.dropout.register_forward_hook(hook)
    for s in range(steps):
        data = torch.randint(0, V, (B, T + 1), generator=gen, dtype=torch.int64)
        x = data[:, :T].transpose(1, 0)          # non-contiguous [T,B] view like main.py:71
        y = data[:, 1:T + 1].transpose(1, 0)
        masks.clear()
        model.zero_grad()
        states = model.detach(states)
        scores, states = model(x, states)
        loss = ref_nll_loss(scores, y)
        loss.backward()
        grads = {k: _np(p.grad) for k, p in model.named_parameters()}
        with torch.no_grad():
            norm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
            for p in model.parameters():
                p -= lr * p.grad
        pre = f"s{s}/"
        out[pre + "x"] = _np(x)
        out[pre + "y"] = _np(y)
        out[pre + "loss"] = np.array(loss.item(), dtype=np.float64)
        out[pre + "norm"] = np.array(float(norm), dtype=np.float64)
        for mi, m in enumerate(masks):
            out[pre + f"mask/{mi}"] = np.packbits(m.reshape(-1))
        for li, (h, c) in enumerate(states):
            out[pre + f"h/{li}"] = _np(h).reshape(B, H)
            out[pre + f"c/{li}"] = _np(c).reshape(B, H)
        if store_full:
            out[pre + "scores"] = _np(scores)
            for k, g in grads.items():
                out[pre + "grad/" + k] = g
            for k, v in model.state_dict().items():
                out[pre + "param/" + k] = _np(v)
        else:
            sc = _np(scores)
            out[pre + "scores_rows"] = sc[:: max(1, sc.shape[0] // 8)][:, :64].copy()
            out[pre + "scores_sum"] = np.array([sc.astype(np.float64).sum(),
                                                np.abs(sc.astype(np.float64)).sum()])
            for k, g in grads.items():
                g64 = g.astype(np.float64)
                out[pre + "grad_l2/" + k] = np.array(np.sqrt((g64 * g64).sum()))
                out[pre + "grad_head/" + k] = g.reshape(-1)[:32].copy()
            for k, v in model.state_dict().items():      # parameters after the update of main.py:116-117
                a = _np(v).astype(np.float64)
                out[pre + "param_sum/" + k] = np.array([a.sum(), np.abs(a).sum(), (a * a).sum()])
                out[pre + "param_head/" + k] = _np(v).reshape(-1)[:32].copy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def minibatch_case():
    out = {}
    cases = [(103, 4, 5), (100, 4, 5), (101, 4, 5), (64, 8, 7), (41, 20, 20), (7, 3, 2), (400, 20, 1)]
    for ci, (n, bs, sl) in enumerate(cases):
        data = np.arange(n).reshape(-1, 1) * 3 % 17
        ds = ref_minibatch(data, bs, sl)
        out[f"c{ci}/args"] = np.array([n, bs, sl])
        out[f"c{ci}/data"] = data
        out[f"c{ci}/n"] = np.array(len(ds))
        for bi, (x, y) in enumerate(ds):
            out[f"c{ci}/x{bi}"] = _np(x)
            out[f"c{ci}/y{bi}"] = _np(y)
    np.savez_compressed(os.path.join(HERE, "minibatch.npz"), **out)
    print("minibatch ok")


def perplexity_case():
    """main.py:86-95 on a slice of the real PTB validation text (token ids only)."""
    V, H, L, T, B = 10000, 8, 2, 7, 5
    with open(os.path.join(REF, "data", "ptb.train.txt")) as f:
        trn = f.read()[1:].split(" ")
    with open(os.path.join(REF, "data", "ptb.valid.txt")) as f:
        vld = f.read()[1:].split(" ")
    words = sorted(set(trn))
    assert len(words) == V
    w2i = {w: i for i, w in enumerate(words)}
    ids = np.array([w2i[w] for w in vld[:1500]]).reshape(-1, 1)
    torch.manual_seed(5)
    model = Model(V, H, L, 0.0, 0.1, "pytorch")
    model.eval()
    _args.batch_size = B
    ds = ref_minibatch(ids, B, T)
    ppl = ref_perplexity(ds, model)
    out = {"meta": np.array([V, H, L, T, B]), "ids": ids.astype(np.int32),
           "ppl": np.array(float(ppl), dtype=np.float64), "n_batches": np.array(len(ds))}
    # only embedding / fc rows that are touched matter, but V*H is small here: store all (fp16-exact not needed)
    for k, v in model.state_dict().items():
        out["param/" + k] = _np(v)
    # ensemble (ensemble.py:97-109) on the first batch with a second model
    torch.manual_seed(6)
    model2 = Model(V, H, L, 0.0, 0.1, "pytorch")
    model2.eval()
    for k, v in model2.state_dict().items():
        out["param2/" + k] = _np(v)
    x, y = ds[0]
    with torch.no_grad():
        s1, _ = model(x, model.state_init(B))
        s2, _ = model2(x, model2.state_init(B))
        out["ens_loss"] = np.array(ref_ensemble_nll_loss([s1, s2], y).item(), dtype=np.float64)
    path = os.path.join(HERE, "perplexity_ptb_slice.npz")
    np.savez_compressed(path, **out)
    print(f"perplexity: ppl={ppl:.4f}  {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    torch.set_num_threads(4)
    only = set(sys.argv[1:])                     # optional: mint only the named fixtures

    def case(name, *a, **kw):
        if not only or name in only:
            step_case(name, *a, **kw)

    #    name               V    H  L  T  B  type      p    winit seed lr  clip
    case("tiny_pytorch",    37, 16, 2, 5, 3, "pytorch", 0.0, 0.3, 11, 1.0, 0.25)
    case("tiny_custom",     37, 16, 2, 5, 3, "custom",  0.0, 0.3, 12, 1.0, 0.25)
    case("tiny_dropout",    41, 24, 2, 6, 4, "pytorch", 0.5, 0.3, 13, 0.5, 5.0, steps=2)
    case("tiny_carry3",     29, 20, 3, 4, 2, "pytorch", 0.0, 0.2, 14, 1.0, 10.0, steps=3, zero_state=True)
    case("edge_T1_B1_L1",   17,  8, 1, 1, 1, "pytorch", 0.0, 0.5, 15, 1.0, 1.0)
    case("odd_H40_custom_dropout", 53, 40, 2, 3, 5, "custom", 0.65, 0.2, 16, 1.0, 2.0)
    case("mid_H72",        150,  72, 2, 20, 20, "pytorch", 0.0, 0.1, 17, 1.0, 5.0, zero_state=True)
    # BASELINE.json configs[0..2] at their exact shapes (README.md:20-27 recipes, V=10000): summaries only, weights
    # re-derived from the seed, the reference's dropout masks stored bit-packed
    case("small_cfg_summary", 10000, 200, 2, 20, 20, "pytorch", 0.0, 0.1, 1, 1.0, 5.0,
         steps=2, zero_state=True, store_full=False)
    case("medium_cfg_summary", 10000, 650, 2, 35, 20, "pytorch", 0.5, 0.05, 1, 1.0, 5.0,
         steps=2, zero_state=True, store_full=False)
    case("large_cfg_summary", 10000, 1500, 2, 35, 20, "pytorch", 0.65, 0.04, 1, 1.0, 10.0,
         steps=2, zero_state=True, store_full=False)
    if not only or "minibatch" in only:
        minibatch_case()
    if not only or "perplexity_ptb_slice" in only:
        perplexity_case()
