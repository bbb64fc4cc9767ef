"""Host logic of the sharded ensemble (CPU): target-probability averaging == ensemble.py:97-126,
against the reference fixture, and the 2-rank gather over gloo."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lstm_lm_oracle as O
from tests._golden import GOLDEN
from zaremba_b200 import ensemble as E


def _fixture():
    z = np.load(os.path.join(GOLDEN, "perplexity_ptb_slice.npz"))
    V, H, L, T, B = [int(v) for v in z["meta"]]
    p1 = {k[len("param/"):]: z[k] for k in z.files if k.startswith("param/")}
    p2 = {k[len("param2/"):]: z[k] for k in z.files if k.startswith("param2/")}
    return z, (V, H, L, T, B), p1, p2


def test_target_prob_mean_equals_reference_ensemble_loss():
    z, (V, H, L, T, B), p1, p2 = _fixture()
    ds = O.minibatch(z["ids"], B, T)
    x, y = ds[0]
    probs = []
    for p in (p1, p2):
        sc, _, _ = O.model_fwd(p, x, O.zero_states(L, B, H), L)
        probs.append(torch.tensor(O.target_probs(sc, y)))
    probs = torch.stack(probs)
    # one batch: exp(mean NLL) == exp(loss / B) of ensemble.py:97-109
    ppl = E.ensemble_perplexity_from_probs(probs, [x.size])
    want = float(np.exp(float(z["ens_loss"]) / B))
    assert abs(ppl - want) < 2e-5 * want
    run = E.running_ensemble_perplexities(probs, [x.size])
    assert len(run) == 2 and abs(run[1] - ppl) < 1e-12
    single = float(np.exp(O.nll_loss(O.model_fwd(p1, x, O.zero_states(L, B, H), L)[0], y) / B))
    assert abs(run[0] - single) < 2e-5 * single


def test_round_robin_placement():
    assert E.models_of_rank(10, 0, 8) == [0, 8] and E.models_of_rank(10, 7, 8) == [7] and E.models_of_rank(2, 3, 8) == []
    assert sorted(sum((E.models_of_rank(10, r, 8) for r in range(8)), [])) == list(range(10))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_models, n_tok = 3, 11
    local = {m: torch.full((n_tok,), 0.1 * (m + 1)) for m in E.models_of_rank(n_models, rank, world)}
    full = E.gather_probs(local, n_models)
    q.put((rank, full.numpy()))
    dist.destroy_process_group()


def test_gather_over_gloo_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([np.full(11, 0.1 * (m + 1), dtype=np.float32) for m in range(3)])
    for r in (0, 1):
        np.testing.assert_allclose(got[r], want, rtol=1e-6)
