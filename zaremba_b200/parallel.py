"""Data-parallel plumbing (host side): one process per GPU, batch rows sharded across ranks.

The reference is single-device.  Its batch rows are independent token streams
(`minibatch` reshapes the corpus to [B, -1], main.py:63-66) and the loss is summed over rows
(main.py:82-84), so the path shards on the batch dimension with exactly one collective per
step: SUM of the flat gradient buffer, after backward and before the global-norm clip
(main.py:113-115).  Every rank then applies the same update to its replica.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend="nccl"):
    """torchrun contract: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* in the environment."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend)
    return rank, local, world


def shard_rows(data, batch_per_rank, rank, world):
    """View the corpus as the global batch [B*world, -1] like main.py:63-66 does for
    `--batch_size B*world`, and return the rows this rank owns (rank*B .. rank*B+B-1) as a
    token column that `minibatch(., B, T)` re-batches into the rank's [T,B] windows."""
    data = np.asarray(data).reshape(-1)
    gb = batch_per_rank * world
    width = data.shape[0] // gb
    rows = data[: width * gb].reshape(gb, width)
    return rows[rank * batch_per_rank:(rank + 1) * batch_per_rank].reshape(-1, 1)


def allreduce_sum_(flat: torch.Tensor, group=None):
    """The one collective of the path: in-place SUM over ranks of the flat gradient buffer."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat
