#!/usr/bin/env python
"""main.py's job (train + validate + test on PTB) driven through the fused `zaremba_b200.Trainer`.

    python tools/train_ptb.py --data /path/to/reference/data --hidden_size 1500 --dropout 0.65 --winit 0.04 \\
        --total_epochs 55 --factor_epoch 14 --factor 1.15 --max_grad_norm 10
    torchrun --nproc-per-node 8 tools/train_ptb.py ...        # data parallel, batch_size rows per GPU

Same flags, data handling (main.py:44-74), LR schedule (main.py:105-106) and log lines (main.py:118-132) as the
reference; the step itself is one library call instead of ~30 eager launches.  (To run the UNMODIFIED main.py on
the drop-in Model instead, see INTEGRATION.md section A.)
"""
import argparse, os, sys, timeit
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import zaremba_b200
from zaremba_b200 import parallel

ap = argparse.ArgumentParser()
ap.add_argument("--data", default="./data")
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
args = ap.parse_args()


def data_init(root):                                   # main.py:44-59
    def read(name):
        with open(os.path.join(root, name)) as f:
            return f.read()[1:].split(" ")
    trn, vld, tst = read("ptb.train.txt"), read("ptb.valid.txt"), read("ptb.test.txt")
    words = sorted(set(trn))
    w2i = {w: i for i, w in enumerate(words)}
    enc = lambda toks: np.array([w2i[w] for w in toks]).reshape(-1, 1)
    return enc(trn), enc(vld), enc(tst), len(words)


rank, local, world = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
trn, vld, tst, vocab = data_init(args.data)
B, T = args.batch_size, args.seq_length
trn_b = zaremba_b200.minibatch(parallel.shard_rows(trn, B, rank, world), B, T)
vld_b = zaremba_b200.minibatch(vld, B, T)
tst_b = zaremba_b200.minibatch(tst, B, T)
torch.manual_seed(args.seed)
model = zaremba_b200.Model(vocab, args.hidden_size, args.layer_num, args.dropout, args.winit).to(f"cuda:{local}")
tr = zaremba_b200.Trainer(model, B, T)
lr, tic, words_seen = args.learning_rate, timeit.default_timer(), 0
for epoch in range(args.total_epochs):
    tr.reset_states()
    model.train()
    if epoch > args.factor_epoch:
        lr = lr / args.factor
    for i, (x, y) in enumerate(trn_b):
        words_seen += x.numel() * world
        if i % max(1, len(trn_b) // 10) == 0:
            loss, norm = tr.train_step_host(x, y, lr, args.max_grad_norm)
            if rank == 0:
                toc = timeit.default_timer()
                print("batch no = {:d} / {:d}, train loss = {:.3f}, wps = {:d}, dw.norm() = {:.3f}, lr = {:.3f}, "
                      "since beginning = {:d} mins, cuda memory = {:.3f} GBs".format(
                          i, len(trn_b), loss / B, round(words_seen / (toc - tic)), norm, lr, round((toc - tic) / 60),
                          torch.cuda.max_memory_allocated() / 1024 ** 3), flush=True)
        else:
            tr.train_step(x.contiguous().cuda(non_blocking=True), y.contiguous().cuda(non_blocking=True), lr,
                          args.max_grad_norm)
    model.eval()
    val = tr.perplexity(vld_b)
    if rank == 0:
        print("Epoch : {:d} || Validation set perplexity : {:.3f}".format(epoch + 1, val))
        print("*************************************************\n", flush=True)
tst_ppl = tr.perplexity(tst_b)
if rank == 0:
    print("Test set perplexity : {:.3f}".format(tst_ppl))
    print("Training is over.")
