"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the CUDA data-parallel step through BOTH gradient
transports (`ce` = copy engines over NVLink peer memory, `nccl` = one all-reduce) against a single process at batch
2B, with dropout ON, and the sharded ensemble (one model per rank) against the reference fixture.

Semantics under test: rows of the global batch are independent streams (main.py:63-66), the loss is
summed over the batch (main.py:82-84), so ranks SUM gradients, clip on the global norm and apply the same update.
Each rank draws its own dropout flags (the rank is folded into the Philox key, `Trainer.seed`); the single-process
run replays the concatenation of the two ranks' masks, fetched through zrb_dropout_mask.
Tolerance: 2e-3 of the parameter scale after 3 steps at lr=1 with an active clip (fp16-operand GEMMs at batch B vs 2B
round differently); replicas must be BIT-identical across ranks.
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests._golden import GOLDEN

pytestmark = pytest.mark.gpu

V, H, L, T, B, P_DROP, STEPS = 1000, 256, 2, 12, 8, 0.5, 3


def _need_two():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def _guarded(fn, rank, world, port, q, *args):
    """Worker entry: an exception in a rank is reported through the queue instead of leaving the parent to time out."""
    try:
        fn(rank, world, port, q, *args)
    except BaseException as e:           # noqa: BLE001 -- report everything, the parent re-raises
        import traceback
        q.put((rank, {"error": "".join(traceback.format_exception(type(e), e, e.__traceback__))[-3000:]}))


def _spawn(fn, world, *args):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_guarded, args=(fn, r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    try:
        for _ in range(world):
            r, res = q.get(timeout=150)
            out[r] = res
            assert "error" not in res, f"rank {r}: {res['error']}"
    finally:
        for p in procs:
            p.join(5 if len(out) < world else 60)
            if p.is_alive():
                p.kill()                 # exactly the processes this test started
    return out


def _dp_worker(rank, world, port, q, transport):
    import ctypes as C
    import zaremba_b200
    from zaremba_b200 import _lib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), ZRB_DP_TRANSPORT=transport)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    data = torch.randint(0, V, (B * world, STEPS * T + 1), generator=g)
    torch.manual_seed(7)
    m = zaremba_b200.Model(V, H, L, P_DROP, 0.1).to(dev)
    m.train()
    tr = zaremba_b200.Trainer(m, B, T)
    assert tr.transport == transport
    rows = slice(rank * B, (rank + 1) * B)
    seeds, out = [], []
    for i in range(STEPS):
        x = data[rows, i * T:(i + 1) * T].t().contiguous().to(dev)
        y = data[rows, i * T + 1:(i + 1) * T + 1].t().contiguous().to(dev)
        seeds.append((tr.seed, tr.step))
        loss, norm = tr.train_step(x, y, 1.0, 0.25)
        out.append((loss.item(), norm.item()))
    dp_p = tr.flat_p.clone()
    # replicas identical?
    bits = dp_p.view(torch.int32).to(torch.int64)
    chk = torch.stack([bits.sum(), (bits * (torch.arange(bits.numel(), device=dev) % 8191 + 1)).sum()])
    hi, lo = chk.clone(), chk.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    identical = bool((hi == lo).all().item())
    # every rank's masks for every step and site -> rank 0
    n = T * B * H
    masks = torch.empty(STEPS, L + 1, n, dtype=torch.uint8, device=dev)
    for i, (seed, step) in enumerate(seeds):
        for site in range(L + 1):
            _lib.check(lib.zrb_dropout_mask(seed, step, site, n, P_DROP, _lib.ptr(masks[i, site]), None))
    allm = [torch.empty_like(masks) for _ in range(world)]
    dist.all_gather(allm, masks)
    losses = torch.tensor([o[0] for o in out], device=dev, dtype=torch.float64)
    dist.all_reduce(losses)                                   # the loss is batch-summed: ranks add up
    res = {"identical": identical, "norms": [o[1] for o in out], "loss_sum": losses.tolist()}
    different_masks = not torch.equal(allm[0], allm[1])
    if rank == 0:
        torch.manual_seed(7)
        m2 = zaremba_b200.Model(V, H, L, P_DROP, 0.1).to(dev)
        m2.train()
        tr2 = zaremba_b200.Trainer(m2, B * world, T, data_parallel=False)
        ref = []
        for i in range(STEPS):
            x = data[:, i * T:(i + 1) * T].t().contiguous().to(dev)
            y = data[:, i * T + 1:(i + 1) * T + 1].t().contiguous().to(dev)
            # global mask [T, 2B, H]: rank r's [T,B,H] block sits at batch rows r*B..
            full = [torch.cat([allm[r][i, site].view(T, B, H) for r in range(world)], dim=1).contiguous()
                    for site in range(L + 1)]
            m2.set_explicit_dropout_masks(full)
            loss, norm = tr2.train_step(x, y, 1.0, 0.25)
            ref.append((loss.item(), norm.item()))
        scale = tr2.flat_p.abs().max().item()
        res.update(err=(dp_p - tr2.flat_p).abs().max().item() / scale, ref=ref, different_masks=different_masks)
    dist.barrier()
    tr.close()
    dist.destroy_process_group()
    q.put((rank, res))


@pytest.mark.parametrize("transport", ["ce", "nccl"])
def test_dp_step_equals_single_process_with_dropout(transport):
    _need_two()
    out = _spawn(_dp_worker, 2, transport)
    r0 = out[0]
    print("DP", transport, r0)
    assert out[0]["identical"] and out[1]["identical"], "replicas diverged across ranks"
    assert r0["different_masks"], "ranks must not share dropout masks"
    assert r0["err"] < 2e-3, r0
    for (l_ref, n_ref), l_dp, n_dp in zip(r0["ref"], r0["loss_sum"], r0["norms"]):
        assert abs(l_dp - l_ref) < 2e-3 * abs(l_ref) and abs(n_dp - n_ref) < 3e-3 * n_ref, r0
    assert r0["norms"][0] > 0.25, "the clip (max_norm 0.25) should be active in this test"


def _ens_worker(rank, world, port, q):
    import zaremba_b200
    from zaremba_b200 import ensemble as E
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    z = np.load(os.path.join(GOLDEN, "perplexity_ptb_slice.npz"))
    Vv, Hh, Ll, Tt, Bb = [int(v) for v in z["meta"]]
    ds = zaremba_b200.minibatch(z["ids"], Bb, Tt)[:1]          # the fixture's ensemble loss is on the first batch
    local = {}
    for mi in E.models_of_rank(2, rank, world):
        pre = "param/" if mi == 0 else "param2/"
        m = zaremba_b200.Model(Vv, Hh, Ll, 0.0, 0.1)
        m.load_state_dict({k[len(pre):]: torch.tensor(z[k]) for k in z.files if k.startswith(pre)})
        m = m.to(dev).eval()
        tr = zaremba_b200.Trainer(m, Bb, Tt, data_parallel=False)
        local[mi], counts = E.target_prob_vector(tr, ds)
    full = E.gather_probs(local, 2)
    ppl = E.ensemble_perplexity_from_probs(full, counts)
    dist.destroy_process_group()
    q.put((rank, {"ppl": ppl, "want": float(np.exp(float(z["ens_loss"]) / Bb)), "models": sorted(local)}))


def test_sharded_ensemble_two_models_two_ranks():
    """ensemble.py:97-109 with model m on rank m: exp(ens_loss / B) of the reference fixture from the gathered
    target probabilities."""
    _need_two()
    out = _spawn(_ens_worker, 2)
    assert out[0]["models"] == [0] and out[1]["models"] == [1]
    for r in (0, 1):
        assert abs(out[r]["ppl"] - out[r]["want"]) < 2e-3 * out[r]["want"], out[r]
