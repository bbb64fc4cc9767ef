"""Ensemble evaluation sharded one-model-per-GPU (BASELINE.json configs[4]; ensemble.py:97-126).

The reference evaluates every model on every batch in one process, stacks the full `[N,V]` probability
tensors, averages them and indexes the target (`ensemble.py:100-106`).  Indexing commutes with the mean,
so all a model has to contribute is its probability OF THE TARGET TOKEN per position:

    NLL_ens(n) = -log( (1/M) * sum_m p_m(y_n) )

Each rank therefore runs `Trainer.eval_step(..., want_probs=True)` over the evaluation set for the models
it hosts (carrying each model's own (h, c), as `ensemble_perplexity` does), and only `[n_tokens]` fp32
vectors cross NVLink (73 k floats for PTB valid) -- replicas only, no gradient exchange.
The reference reports the running ensemble after each model is added (`ensemble.py:177-180`); with the
per-model vectors gathered that is a prefix mean.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def target_prob_vector(trainer, batches):
    """softmax(scores)[n, y_n] for every position of `batches`, in order, for one model.
    Also returns the per-batch token counts so the caller can rebuild main.py's mean of batch means."""
    trainer.reset_states()
    dev = trainer.dev
    out, counts = [], []
    for x, y in batches:
        xd = x.to(dev).contiguous()
        yd = y.to(dev).contiguous()
        _, tp = trainer.eval_step(xd, yd, want_probs=True)
        out.append(tp.clone())
        counts.append(x.numel())
    return torch.cat(out), counts


def ensemble_perplexity_from_probs(probs, counts):
    """probs [M, n_tokens] target probabilities of M models; counts = tokens per batch.
    Returns exp(mean over batches of the batch-mean ensemble NLL) -- exactly what
    `ensemble_perplexity` computes (ensemble.py:111-126: mean of per-batch `loss/batch_size`)."""
    pbar = probs.double().mean(0)
    nll = -torch.log(pbar)
    losses, off = [], 0
    for c in counts:
        losses.append(nll[off:off + c].mean())
        off += c
    return math.exp(torch.stack(losses).mean().item())


def running_ensemble_perplexities(probs, counts):
    """[ppl with model 1, ppl with models 1-2, ...] like the reference's report after each model."""
    return [ensemble_perplexity_from_probs(probs[: m + 1], counts) for m in range(probs.shape[0])]


def models_of_rank(n_models, rank, world):
    """Round-robin placement: model m lives on rank m % world (10 Large models on 8 GPUs -> two waves)."""
    return [m for m in range(n_models) if m % world == rank]


def gather_probs(local, n_models, group=None):
    """local: {model_index: [n_tokens] tensor} on this rank.  Returns [n_models, n_tokens] on every rank
    (one all_reduce of a zero-padded matrix: each row is owned by exactly one rank)."""
    any_vec = next(iter(local.values())) if local else None
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return torch.stack([local[m] for m in range(n_models)])
    if any_vec is not None:
        dev = any_vec.device
    else:   # a rank that hosts no model still joins the collectives: NCCL needs a CUDA tensor
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    n_tok = torch.tensor([0 if any_vec is None else any_vec.numel()], device=dev)
    dist.all_reduce(n_tok, op=dist.ReduceOp.MAX, group=group)
    full = torch.zeros(n_models, int(n_tok.item()), device=dev, dtype=torch.float32)
    for m, v in local.items():
        full[m] = v
    dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return full
