"""world_size-2 gloo test of the data-parallel semantics (CPU): rows sharded across ranks,
ONE sum all-reduce of the flat gradients, then every rank clips on the global norm and applies
SGD == a single process at batch_size 2B.  Per-rank arithmetic is the oracle's;
what is under test is the package's sharding (`parallel.shard_rows`, `minibatch`) and reduction
helper (`parallel.allreduce_sum_`)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lstm_lm_oracle as O

V, H, L, T, B = 31, 10, 2, 4, 3


def _corpus():
    return np.random.default_rng(0).integers(0, V, size=2 * B * (3 * T + 1) + 5)


def _worker(rank, world, port, q):
    import zaremba_b200
    from zaremba_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names = O.param_names(L)
    params = O.init_params(V, H, L, 0.3, 7, np.float64)
    mine = parallel.shard_rows(_corpus(), B, rank, world)
    batches = zaremba_b200.minibatch(mine, B, T)
    states = O.zero_states(L, B, H, np.float64)
    for x, y in batches[:2]:
        sc, states, cache = O.model_fwd(params, x.numpy(), states, L)
        grads = O.model_bwd(params, cache, O.nll_loss_bwd(sc, y.numpy()), L)
        flat = torch.tensor(np.concatenate([grads[n].reshape(-1) for n in names]))
        parallel.allreduce_sum_(flat)
        off = 0
        for n in names:
            k = grads[n].size
            grads[n] = flat[off:off + k].numpy().reshape(grads[n].shape).copy()
            off += k
        O.clip_sgd(params, grads, 1.0, 0.5, names)
    q.put((rank, {n: params[n] for n in names}))
    dist.destroy_process_group()


def test_two_rank_dp_equals_single_process_double_batch():
    import zaremba_b200
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process, batch 2B, same corpus
    names = O.param_names(L)
    params = O.init_params(V, H, L, 0.3, 7, np.float64)
    batches = zaremba_b200.minibatch(_corpus().reshape(-1, 1), 2 * B, T)
    states = O.zero_states(L, 2 * B, H, np.float64)
    for x, y in batches[:2]:
        sc, states, cache = O.model_fwd(params, x.numpy(), states, L)
        # loss = mean_n(-log p) * batch: with 2B rows the per-row gradient weight is the same 1/T
        grads = O.model_bwd(params, cache, O.nll_loss_bwd(sc, y.numpy()), L)
        O.clip_sgd(params, grads, 1.0, 0.5, names)
    for r in (0, 1):
        for n in names:
            np.testing.assert_allclose(got[r][n], params[n], rtol=1e-10, atol=1e-12)
