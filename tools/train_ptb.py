#!/usr/bin/env python
"""main.py's job (train + validate + test on Penn Treebank) driven through the fused `zaremba_b200.Trainer`,
or -- `--impl cudnn` -- through the reference's own `--lstm_type pytorch` sequence of torch calls on the GPU
(oracle/torch_port.py: cuDNN nn.LSTM, eager loss, clip_grad_norm_, per-parameter SGD), same seed, same data,
same schedule: the "matched valid perplexity" half of BASELINE.json's target.

    python tools/train_ptb.py --recipe large                       # README.md:26 on the committed id fixture
    python tools/train_ptb.py --recipe small --impl cudnn --json gpurun_out/ptb_small_cudnn.json
    python tools/train_ptb.py --data /root/reference/data --hidden_size 650 ...   # from the text files
    torchrun --nproc-per-node 8 tools/train_ptb.py --recipe large  # data parallel, batch_size rows per GPU

Same flags, data handling (main.py:44-74), LR schedule (main.py:105-106) and log lines (main.py:118-132) as the
reference; the step itself is one library call instead of ~30 eager launches.  (To run the UNMODIFIED main.py on
the drop-in Model instead, see INTEGRATION.md section A.)  Token ids come from `tests/golden/ptb_ids.npz`
(minted from the reference's text by tests/golden/make_ptb_ids.py with the reference's vocabulary rule) unless
`--data` names a directory with the ptb.*.txt files.  `--impl cudnn` is baseline/test infrastructure: it is the
only mode that imports `oracle/`.
"""
import argparse, json, math, os, sys, time, timeit
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

RECIPES = {   # README.md:20-27
    "small": dict(hidden_size=200, dropout=0.0, winit=0.1, seq_length=20, total_epochs=13, factor_epoch=4, factor=2.0,
                  max_grad_norm=5.0),
    "medium": dict(hidden_size=650, dropout=0.5, winit=0.05, seq_length=35, total_epochs=39, factor_epoch=6,
                   factor=1.2, max_grad_norm=5.0),
    "large": dict(hidden_size=1500, dropout=0.65, winit=0.04, seq_length=35, total_epochs=55, factor_epoch=14,
                  factor=1.15, max_grad_norm=10.0),
}

ap = argparse.ArgumentParser()
ap.add_argument("--recipe", choices=sorted(RECIPES), default=None, help="one of the README's three single-model recipes")
ap.add_argument("--impl", choices=["ours", "cudnn"], default="ours")
ap.add_argument("--data", default=None, help="directory with ptb.{train,valid,test}.txt (default: the id fixture)")
ap.add_argument("--ids", default=os.path.join(ROOT, "tests", "golden", "ptb_ids.npz"))
ap.add_argument("--layer_num", type=int, default=2)
ap.add_argument("--hidden_size", type=int, default=650)
ap.add_argument("--dropout", type=float, default=0.5)
ap.add_argument("--winit", type=float, default=0.05)
ap.add_argument("--batch_size", type=int, default=20)
ap.add_argument("--seq_length", type=int, default=35)
ap.add_argument("--learning_rate", type=float, default=1)
ap.add_argument("--total_epochs", type=int, default=39)
ap.add_argument("--factor_epoch", type=int, default=6)
ap.add_argument("--factor", type=float, default=1.2)
ap.add_argument("--max_grad_norm", type=float, default=5)
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--lazy_update", action="store_true",
                help="Trainer(lazy_update=True): upper-layer / fc weight updates run beside the next step's forward")
ap.add_argument("--eval_batch_size", type=int, default=None,
                help="batch size of the validation / test sweeps (default: --batch_size, like main.py)")
ap.add_argument("--epochs", type=int, default=None, help="stop after this many epochs (schedule unchanged)")
ap.add_argument("--json", default=None, help="write per-epoch validation perplexities, test perplexity, timing here")
ap.add_argument("--save", default=None, help="save the trained state_dict (reference key names) here")
args = ap.parse_args()
if args.recipe:
    for k, v in RECIPES[args.recipe].items():
        setattr(args, k, v)


def data_init_text(root):                              # main.py:44-59
    def read(name):
        with open(os.path.join(root, name)) as f:
            return f.read()[1:].split(" ")
    trn, vld, tst = read("ptb.train.txt"), read("ptb.valid.txt"), read("ptb.test.txt")
    words = sorted(set(trn))
    w2i = {w: i for i, w in enumerate(words)}
    enc = lambda toks: np.array([w2i[w] for w in toks]).reshape(-1, 1)
    return enc(trn), enc(vld), enc(tst), len(words)


def data_init_ids(path):
    d = np.load(path)
    col = lambda a: a.astype(np.int64).reshape(-1, 1)
    return col(d["train"]), col(d["valid"]), col(d["test"]), int(d["vocab_size"])


import zaremba_b200
from zaremba_b200 import parallel

rank, local, world = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
trn, vld, tst, vocab = data_init_text(args.data) if args.data else data_init_ids(args.ids)
B, T = args.batch_size, args.seq_length
trn_b = zaremba_b200.minibatch(parallel.shard_rows(trn, B, rank, world), B, T)
EB = args.eval_batch_size or B
vld_b = zaremba_b200.minibatch(vld, EB, T)
tst_b = zaremba_b200.minibatch(tst, EB, T)
torch.manual_seed(args.seed)

if args.impl == "ours":
    model = zaremba_b200.Model(vocab, args.hidden_size, args.layer_num, args.dropout, args.winit).to(dev)
    tr = zaremba_b200.Trainer(model, B, T, lazy_update=args.lazy_update)
    # the corpus is staged on the device once: 3 x [n_batches, T, B] int64
    trn_x = torch.stack([x for x, _ in trn_b]).contiguous().to(dev)
    trn_y = torch.stack([y for _, y in trn_b]).contiguous().to(dev)

    def train_epoch(lr, log):
        tr.reset_states()
        model.train()
        every = max(1, len(trn_b) // 10)
        for i in range(len(trn_b)):
            loss, norm = tr.train_step(trn_x[i], trn_y[i], lr, args.max_grad_norm)
            if i % every == 0:
                log(i, loss.item(), norm.item())

    def perplexity(batches):
        model.eval()
        return tr.perplexity(batches)

    def state_dict():
        tr.flush()
        return {k: v.detach().cpu() for k, v in model.state_dict().items()}
else:
    if world > 1:
        raise SystemExit("--impl cudnn is the reference's single-device path")
    from oracle import torch_port as P
    model = P.TorchLstmLm(vocab, args.hidden_size, args.layer_num, args.dropout, args.winit).to(dev)
    trn_d = [(x.to(dev), y.to(dev)) for x, y in trn_b]

    def train_epoch(lr, log):
        model.train()
        states = model.zero_state(B)
        every = max(1, len(trn_b) // 10)
        for i, (x, y) in enumerate(trn_d):
            loss, norm, states = P.train_step(model, x, y, states, lr, args.max_grad_norm)
            if i % every == 0:
                log(i, loss.item(), float(norm))

    def perplexity(batches):                           # main.py:86-95
        model.eval()
        with torch.no_grad():
            losses, states = [], model.zero_state(EB)
            for x, y in batches:
                logits, states = model(x.to(dev), states)
                losses.append(P.softmax_nll_times_batch(logits, y.to(dev)).item() / EB)
        return float(np.exp(np.mean(losses)))

    def state_dict():
        return {k: v.detach().cpu() for k, v in model.reference_state_dict().items()}

lr, tic, words_seen = args.learning_rate, timeit.default_timer(), 0
val_curve, epoch_secs = [], []
n_epochs = args.total_epochs if args.epochs is None else min(args.epochs, args.total_epochs)
for epoch in range(n_epochs):
    if epoch > args.factor_epoch:                      # main.py:105-106
        lr = lr / args.factor
    e_words = [0]

    def log(i, loss, norm, epoch=epoch):
        if loss != loss:                               # NaN: the run is dead, do not burn the remaining epochs
            if rank == 0:
                print(f"NON-FINITE train loss at epoch {epoch + 1}, batch {i}: aborting (seed {args.seed})", flush=True)
                if args.json:
                    json.dump({"impl": args.impl, "recipe": args.recipe, "seed": args.seed, "diverged": True,
                               "epoch": epoch + 1, "batch": i, "valid_ppl_per_epoch": [round(v, 3) for v in val_curve]},
                              open(args.json, "w"), indent=1)
            sys.exit(3)
        if rank == 0:
            toc = timeit.default_timer()
            seen = words_seen + (i + 1) * T * B * world
            print("batch no = {:d} / {:d}, train loss = {:.3f}, wps = {:d}, dw.norm() = {:.3f}, lr = {:.3f}, "
                  "since beginning = {:d} mins, cuda memory = {:.3f} GBs".format(
                      i, len(trn_b), loss / B, round(seen / (toc - tic)), norm, lr, round((toc - tic) / 60),
                      torch.cuda.max_memory_allocated() / 1024 ** 3), flush=True)

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    train_epoch(lr, log)
    torch.cuda.synchronize()
    epoch_secs.append(time.perf_counter() - t0)
    words_seen += len(trn_b) * T * B * world
    val = perplexity(vld_b)
    val_curve.append(val)
    if rank == 0:
        print("Epoch : {:d} || Validation set perplexity : {:.3f}".format(epoch + 1, val))
        print("*************************************************\n", flush=True)
tst_ppl = perplexity(tst_b)
if rank == 0:
    print("Test set perplexity : {:.3f}".format(tst_ppl))
    print("Training is over.")
    if args.json:
        steps = len(trn_b)
        out = {"impl": args.impl, "recipe": args.recipe, "args": {k: v for k, v in vars(args).items()
                                                                 if k not in ("json", "save", "data", "ids")},
               "world": world, "vocab": vocab, "steps_per_epoch": steps, "epochs_run": n_epochs,
               "valid_ppl_per_epoch": [round(v, 3) for v in val_curve], "test_ppl": round(tst_ppl, 3),
               "train_seconds_per_epoch_median": float(np.median(epoch_secs)),
               "train_tokens_per_s_median_epoch": steps * T * B * world / float(np.median(epoch_secs)),
               "total_wall_s": timeit.default_timer() - tic, "gpu": torch.cuda.get_device_name(0),
               "data": os.path.basename(args.data or args.ids)}
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        json.dump(out, open(args.json, "w"), indent=1)
    if args.save:
        torch.save(state_dict(), args.save)
