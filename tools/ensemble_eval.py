#!/usr/bin/env python
"""ensemble.py's job (train M models, report the running model-averaged perplexity) sharded ONE MODEL PER GPU.

    torchrun --nproc-per-node 8 tools/ensemble_eval.py --recipe large --ensemble_num 10 --epochs 6 \\
        --json gpurun_out/ensemble_large10.json
    python tools/ensemble_eval.py --recipe small --ensemble_num 2            # README.md:35, one GPU, two waves

BASELINE.json configs[4]; reference: ensemble.py:97-126 (ensemble_nll_loss / ensemble_perplexity) and :166-180 (the
driver).  The reference trains the models one after the other on one device and, after each, re-evaluates ALL
models so far on every batch, stacking M full [N,V] probability tensors.  Here model m lives on rank m % world
(`models_of_rank`), trains alone (replicas only: no gradient exchange, `Trainer(data_parallel=False)`), and
contributes only its probability OF THE TARGET TOKEN per position (`zrb_eval_step(tgt_prob)`): indexing commutes
with the mean over models, so [n_tokens] floats per model and split are all that crosses NVLink
(`gather_probs`), and the reference's "perplexity of the first k models" reports are prefix means.

Each model is initialised from `torch.manual_seed(seed + m)` and sees the whole training corpus, like the
reference's models (which differ only through the global RNG state).  `--epochs` shortens training (the LR schedule is
the recipe's); timing of the evaluation sweep is reported as models*tokens/s over all ranks.
"""
import argparse, json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

RECIPES = {   # README.md:34-41 (same hyper-parameters as the single models)
    "small": dict(hidden_size=200, dropout=0.0, winit=0.1, seq_length=20, total_epochs=13, factor_epoch=4, factor=2.0,
                  max_grad_norm=5.0),
    "medium": dict(hidden_size=650, dropout=0.5, winit=0.05, seq_length=35, total_epochs=39, factor_epoch=6,
                   factor=1.2, max_grad_norm=5.0),
    "large": dict(hidden_size=1500, dropout=0.65, winit=0.04, seq_length=35, total_epochs=55, factor_epoch=14,
                  factor=1.15, max_grad_norm=10.0),
}
ap = argparse.ArgumentParser()
ap.add_argument("--recipe", choices=sorted(RECIPES), default="large")
ap.add_argument("--ensemble_num", type=int, default=10)
ap.add_argument("--epochs", type=int, default=None, help="train each model this many epochs (default: the recipe's)")
ap.add_argument("--batch_size", type=int, default=20)
ap.add_argument("--layer_num", type=int, default=2)
ap.add_argument("--learning_rate", type=float, default=1.0)
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--ids", default=os.path.join(ROOT, "tests", "golden", "ptb_ids.npz"))
ap.add_argument("--json", default=None)
args = ap.parse_args()
R = RECIPES[args.recipe]

import zaremba_b200
from zaremba_b200 import parallel
from zaremba_b200 import ensemble as E

rank, local, world = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
d = np.load(args.ids)
col = lambda a: a.astype(np.int64).reshape(-1, 1)
trn, vld, tst, vocab = col(d["train"]), col(d["valid"]), col(d["test"]), int(d["vocab_size"])
B, T = args.batch_size, R["seq_length"]
trn_b = zaremba_b200.minibatch(trn, B, T)
vld_b = zaremba_b200.minibatch(vld, B, T)
tst_b = zaremba_b200.minibatch(tst, B, T)
trn_x = torch.stack([x for x, _ in trn_b]).contiguous().to(dev)
trn_y = torch.stack([y for _, y in trn_b]).contiguous().to(dev)
n_epochs = R["total_epochs"] if args.epochs is None else min(args.epochs, R["total_epochs"])

mine = E.models_of_rank(args.ensemble_num, rank, world)
local_v, local_t, single, train_s, eval_s, eval_tokens = {}, {}, {}, 0.0, 0.0, 0
counts_v = counts_t = None
for m in mine:
    torch.manual_seed(args.seed + m)
    model = zaremba_b200.Model(vocab, R["hidden_size"], args.layer_num, R["dropout"], R["winit"]).to(dev)
    tr = zaremba_b200.Trainer(model, B, T, data_parallel=False)
    lr = args.learning_rate
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for epoch in range(n_epochs):
        tr.reset_states()
        model.train()
        if epoch > R["factor_epoch"]:
            lr = lr / R["factor"]
        for i in range(len(trn_b)):
            tr.train_step(trn_x[i], trn_y[i], lr, R["max_grad_norm"])
    torch.cuda.synchronize(); train_s += time.perf_counter() - t0
    model.eval()
    t0 = time.perf_counter()
    local_v[m], counts_v = E.target_prob_vector(tr, vld_b)
    local_t[m], counts_t = E.target_prob_vector(tr, tst_b)
    torch.cuda.synchronize(); eval_s += time.perf_counter() - t0
    eval_tokens += int(local_v[m].numel() + local_t[m].numel())
    single[m] = (E.ensemble_perplexity_from_probs(local_v[m][None], counts_v),
                 E.ensemble_perplexity_from_probs(local_t[m][None], counts_t))
    print(f"[rank {rank}] model {m + 1}: valid {single[m][0]:.3f} test {single[m][1]:.3f} "
          f"({n_epochs} epochs, {train_s:.1f} s training so far)", flush=True)
    del tr, model
    torch.cuda.empty_cache()

if counts_v is None:                                   # a rank without a model still joins the collectives
    counts_v, counts_t = [x.numel() for x, _ in vld_b], [x.numel() for x, _ in tst_b]
full_v = E.gather_probs(local_v, args.ensemble_num)
full_t = E.gather_probs(local_t, args.ensemble_num)
stats = torch.tensor([train_s, eval_s, float(eval_tokens)], device=dev, dtype=torch.float64)
if world > 1:
    mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
else:
    mx = sm = stats
if rank == 0:
    run_v = E.running_ensemble_perplexities(full_v, counts_v)
    run_t = E.running_ensemble_perplexities(full_t, counts_t)
    one_v = [E.ensemble_perplexity_from_probs(full_v[m][None], counts_v) for m in range(args.ensemble_num)]
    one_t = [E.ensemble_perplexity_from_probs(full_t[m][None], counts_t) for m in range(args.ensemble_num)]
    for k in range(args.ensemble_num):                 # ensemble.py:177-180
        print("Validation set perplexity of {} averaged models: {:.3f}".format(k + 1, run_v[k]))
        print("Test set perplexity of {} averaged models: {:.3f}\n".format(k + 1, run_t[k]))
    out = {"recipe": args.recipe, "ensemble_num": args.ensemble_num, "world": world, "epochs_per_model": n_epochs,
           "placement": {str(r): E.models_of_rank(args.ensemble_num, r, world) for r in range(world)},
           "single_model_valid_ppl": [round(v, 3) for v in one_v], "single_model_test_ppl": [round(v, 3) for v in one_t],
           "running_ensemble_valid_ppl": [round(v, 3) for v in run_v],
           "running_ensemble_test_ppl": [round(v, 3) for v in run_t],
           "train_seconds_max_over_ranks": mx[0].item(), "eval_seconds_max_over_ranks": mx[1].item(),
           "eval_model_tokens_total": int(sm[2].item()),
           "eval_model_tokens_per_s": sm[2].item() / max(mx[1].item(), 1e-9),
           "bytes_crossing_gpus": int(4 * (full_v.numel() + full_t.numel())), "gpu": torch.cuda.get_device_name(0)}
    print(json.dumps(out))
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        json.dump(out, open(args.json, "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
