"""Watchdog of the persistent recurrence kernels (include/zaremba_b200.h: zrb_check_health; csrc/rec_common.cuh: RecWatch).

Fault injection: ZRB_FAULT_BARRIER_BASE="fwd" / "bwd" makes that kernel's launches expect one grid-barrier arrival more
than the grid delivers, so every CTA waits for a count that never comes -- the lost wake-up the bounded waits exist for.  With the time-out cut
to ~20 ms (ZRB_SPIN_CYCLES) the kernel must (1) terminate instead of hanging or trapping, (2) leave the CUDA context
usable, (3) make the next library call fail with ZRB_E_CUDA naming the wait.  Runs in a subprocess: the switches are read
once per process.
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, time
sys.path.insert(0, %r)
import torch, zaremba_b200
from zaremba_b200 import _lib
V, H, L, T, B = 500, 256, 2, 9, 8
torch.manual_seed(0)
m = zaremba_b200.Model(V, H, L, 0.0, 0.1).cuda(); m.train()
tr = zaremba_b200.Trainer(m, B, T)
d = torch.randint(0, V, (B, T + 1))
x, y = d[:, :T].t().contiguous().cuda(), d[:, 1:].t().contiguous().cuda()
t0 = time.time()
tr.train_step(x, y, 1.0, 5.0)            # the faulty kernel is inside; the call itself is asynchronous
torch.cuda.synchronize()                 # (1) terminates, (2) no CUDA error
dt = time.time() - t0
assert dt < 20, dt
rc = _lib.load().zrb_check_health(tr.ctx)
assert rc != 0, "the watchdog word should be set"
try:
    tr.train_step(x, y, 1.0, 5.0)
    raise SystemExit("second step should have failed")
except RuntimeError as e:
    msg = str(e)
    assert "gave up waiting" in msg and "grid barrier" in msg, msg
z = (torch.ones(1024, device="cuda") * 2).sum().item()   # (2) the CUDA context still works
assert z == 2048.0
print("WATCHDOG_OK", round(dt, 3), msg[:160])
"""


@pytest.mark.parametrize("which", ["fwd", "bwd"])
def test_lost_arrival_is_reported_not_fatal(which):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    env = dict(os.environ, ZRB_FAULT_BARRIER_BASE=which, ZRB_SPIN_CYCLES="40000000")
    r = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "WATCHDOG_OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])


def test_healthy_context_reports_ok():
    import zaremba_b200
    from zaremba_b200 import _lib
    m = zaremba_b200.Model(300, 96, 2, 0.0, 0.1).cuda()
    tr = zaremba_b200.Trainer(m, 4, 5)
    d = torch.randint(0, 300, (4, 6))
    tr.train_step(d[:, :5].t().contiguous().cuda(), d[:, 1:].t().contiguous().cuda(), 1.0, 5.0)
    torch.cuda.synchronize()
    assert _lib.load().zrb_check_health(tr.ctx) == 0
