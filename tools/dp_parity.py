#!/usr/bin/env python
"""2-rank NCCL data-parallel step == single-process step at batch 2B (GPU).  Run under torchrun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import zaremba_b200

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
V, H, L, T, B = 1000, 256, 2, 12, 8
g = torch.Generator().manual_seed(3)
data = torch.randint(0, V, (B * world, 3 * T + 1), generator=g)
def run(model, rows, nb):
    tr = zaremba_b200.Trainer(model, nb, T)
    out = []
    for i in range(3):
        x = data[rows, i * T:(i + 1) * T].t().contiguous().cuda(); y = data[rows, i * T + 1:(i + 1) * T + 1].t().contiguous().cuda()
        loss, norm = tr.train_step(x, y, 1.0, 0.5)
        out.append((loss.item(), norm.item()))
    return out, tr.flat_p.clone()
torch.manual_seed(7)
m = zaremba_b200.Model(V, H, L, 0.0, 0.1).cuda(); m.train()
dp_out, dp_p = run(m, slice(rank * B, (rank + 1) * B), B)
# single process, batch 2B, without the process group's all-reduce
torch.manual_seed(7)
m2 = zaremba_b200.Model(V, H, L, 0.0, 0.1).cuda(); m2.train()
tr2 = zaremba_b200.Trainer(m2, B * world, T); tr2.world = 1
ref = []
for i in range(3):
    x = data[:, i * T:(i + 1) * T].t().contiguous().cuda(); y = data[:, i * T + 1:(i + 1) * T + 1].t().contiguous().cuda()
    loss, norm = tr2.train_step(x, y, 1.0, 0.5); ref.append((loss.item(), norm.item()))
err = (dp_p - tr2.flat_p).abs().max().item(); scale = tr2.flat_p.abs().max().item()
losses = torch.tensor([o[0] for o in dp_out], device="cuda"); dist.all_reduce(losses)   # loss is batch-summed: ranks add up
print(f"rank {rank}: max param diff {err:.3e} (scale {scale:.3e}); dp loss sum {losses.tolist()} vs single {[r[0] for r in ref]}; norms {[o[1] for o in dp_out]} vs {[r[1] for r in ref]}")
assert err < 5e-3 * scale
dist.destroy_process_group()
print("DP_PARITY_OK")
