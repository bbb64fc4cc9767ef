#!/usr/bin/env python
"""bench.py -- tokens/sec of the LSTM-LM train step (main.py:109-117) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config large|medium|small]
                    [--impl ours|reference] [--engine tc|simt]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one synthetic
[T,B] window per GPU: forward, softmax-NLL, backward, (gradient all-reduce when N>1),
global-norm clip, SGD.  Weak scaling: B=20 per GPU, rows of the global batch are
independent token streams (main.py:63-66), so no data-path collective except the one
gradient all-reduce.

  value      whole-job tokens/s, tokens already resident in HBM, CUDA events, max over ranks
  e2e        same step through the public host-buffer call (`Trainer.train_step_host` ->
             zrb_train_step_host): CPU [T,B] int64 views as main.py:71-72 builds them, H2D of
             x,y and D2H of the loss inside the timed region
  roofline   dominant kernel class, timed live with CUDA events on the launching stream in a
             second instrumented pass of the same K steps (zrb_prof_*)
  cpu_baseline  oracle/torch_port.py (the reference's own torch calls) on the host cores
  --impl reference   that CPU port alone, same config / metric (rank 0 only)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: V, H, L, T, B, dropout, winit, lr, max_norm   (README.md:20-27 recipes)
    "small": dict(V=10000, H=200, L=2, T=20, B=20, p=0.0, winit=0.1, lr=1.0, clip=5.0),
    "medium": dict(V=10000, H=650, L=2, T=35, B=20, p=0.5, winit=0.05, lr=1.0, clip=5.0),
    "large": dict(V=10000, H=1500, L=2, T=35, B=20, p=0.65, winit=0.04, lr=1.0, clip=10.0),
}
METRIC = "tokens/sec (and valid perplexity) Large-LSTM 1500h at 1/2/4/8 B200 vs ref CPU"


def workload_name(name, c):
    return (f"{name}: {c['L']}x{c['H']} LSTM LM train step, T={c['T']}, B={c['B']}/GPU, V={c['V']}, "
            f"dropout {c['p']}, synthetic uniform tokens")


def flops_per_token(c):
    """train FLOPs/token = 3 * (16*L*H^2 + 2*H*V)."""
    return 3 * (16 * c["L"] * c["H"] ** 2 + 2 * c["H"] * c["V"])


def class_work(c):
    """Algorithmic work of one step per kernel class: (kind, amount) with FLOPs for the dense
    contractions and bytes for the streaming kernels (DESIGN.md section 'Kernels')."""
    N, H, L, V = c["T"] * c["B"], c["H"], c["L"], c["V"]
    P = 2 * V * H + V + L * (8 * H * H + 8 * H)
    return {
        "gemm_in": ("flop", 8 * N * H * H * L), "rec_fwd": ("flop", 8 * N * H * H * L),
        "proj_fwd": ("flop", 2 * N * H * V), "proj_bwd": ("flop", 4 * N * H * V),
        "rec_bwd": ("flop", 8 * N * H * H * L), "gemm_dx": ("flop", 8 * N * H * H * L),
        "gemm_wgrad": ("flop", 16 * N * H * H * L),
        "softmax": ("byte", 2 * N * V * 4), "clip_sgd": ("byte", 20 * P),
        "embed_fwd": ("byte", 2 * N * H * 4), "embed_bwd": ("byte", V * H * 4 + 2 * N * H * 4),
        "pack": ("byte", 6 * (P - V * H - V)),
    }


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tensor=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([f.strip() for f in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 8 and r[4 + i].lower().startswith("active")
                                                         for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_port_leg(c, budget_s, steps=None, warmup=1):
    """The reference's CPU path (torch port) on the host cores; bounded sample."""
    import torch
    from oracle import torch_port as P
    if steps is None:
        dt1, _, thr = P.time_cpu_train_steps(c["V"], c["H"], c["L"], c["B"], c["T"], c["p"], c["winit"], c["lr"],
                                             c["clip"], steps=1, warmup=1)
        steps = max(3, min(200, int(budget_s / max(dt1, 1e-3))))
    dt, tps, thr = P.time_cpu_train_steps(c["V"], c["H"], c["L"], c["B"], c["T"], c["p"], c["winit"], c["lr"],
                                          c["clip"], steps=steps, warmup=warmup)
    return {"value": tps, "unit": "tokens/s", "cores": thr, "kind": "port",
            "sample": f"{steps} train steps of the same config ({steps * c['T'] * c['B']} tokens) after {warmup} "
                      f"warm-up, torch {torch.__version__} CPU (oneDNN nn.LSTM), {dt * 1e3:.1f} ms/step",
            "host_cpus": os.cpu_count()}, dt


def run_reference(args, c, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: at most ~100 s of CPU work whatever --steps says (0.73 s/step for Large on 64 threads)
    probe, dt1 = cpu_port_leg(c, 0, steps=1, warmup=1)
    run_steps = max(3, min(args.steps, int(100.0 / max(dt1, 1e-3))))
    base, dt = cpu_port_leg(c, 0, steps=run_steps, warmup=min(max(1, args.warmup), 3))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(name, c), "device": "host CPU"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args, c, name):
    import torch
    import torch.distributed as dist
    import zaremba_b200
    from zaremba_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: zaremba_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    V, H, L, T, B = c["V"], c["H"], c["L"], c["T"], c["B"]
    K, W = args.steps, args.warmup
    torch.manual_seed(1)                       # same weights on every rank (replicated parameters)
    model = zaremba_b200.Model(V, H, L, c["p"], c["winit"], engine=args.engine).to(dev)
    model.train()
    tr = zaremba_b200.Trainer(model, B, T)
    # synthetic PTB-shaped tokens: the global batch is [B*world, .]; this rank owns rows rank*B .. rank*B+B-1
    g = torch.Generator().manual_seed(2)
    n_win = K + W
    data = torch.randint(0, V, (B * world, T * n_win + 1), generator=g, dtype=torch.int64)[rank * B:(rank + 1) * B]
    host_batches = [(data[:, i * T:(i + 1) * T].t(), data[:, i * T + 1:(i + 1) * T + 1].t()) for i in range(n_win)]
    dev_batches = [(x.contiguous().to(dev), y.contiguous().to(dev)) for x, y in host_batches]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(fn, batches):
        for x, y in batches[:W]:
            fn(x, y)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.zrb_launch_count()
        t0 = time.perf_counter()
        e0.record()
        for x, y in batches[W:W + K]:
            fn(x, y)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms[0].item(), ms[1].item(), lib.zrb_launch_count() - n0

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    # 1) device-resident tokens: the headline `value`
    dev_ms, _, launches = timed_region(lambda x, y: tr.train_step(x, y, c["lr"], c["clip"]), dev_batches)
    clocks = sampler.stop() if sampler else None
    loss_after = tr.loss.item()
    # 2) end to end through the host-buffer call (wall clock: H2D, step, D2H of the loss each step)
    tr.reset_states()
    _, e2e_wall_ms, _ = timed_region(lambda x, y: tr.train_step_host(x, y, c["lr"], c["clip"]), host_batches)
    # 3) instrumented pass for the per-class roofline
    import ctypes as C
    tr.reset_states()
    _lib.check(lib.zrb_prof_enable(tr.ctx, 1))
    for x, y in dev_batches[W:W + K]:
        tr.train_step(x, y, c["lr"], c["clip"])
    ms_arr = (C.c_float * len(_lib.PROF_CLASSES))()
    cnt_arr = (C.c_int64 * len(_lib.PROF_CLASSES))()
    _lib.check(lib.zrb_prof_read(tr.ctx, ms_arr, cnt_arr))
    _lib.check(lib.zrb_prof_enable(tr.ctx, 0))
    per_class = {n: (ms_arr[i] / K, cnt_arr[i] / K) for i, n in enumerate(_lib.PROF_CLASSES) if cnt_arr[i]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    tokens = T * B * world * K
    peaks = measured_peaks()
    work = class_work(c)
    top = max(per_class, key=lambda n: per_class[n][0]) if per_class else None
    roofline = None
    if top:
        kind, amount = work[top]
        ms_step = per_class[top][0]
        if kind == "flop":
            ach, peak, unit, bound = amount / (ms_step * 1e-3) / 1e12, peaks["tensor"], "TFLOP/s", "tensor"
        else:
            ach, peak, unit, bound = amount / (ms_step * 1e-3) / 1e9, peaks["hbm"], "GB/s", "hbm"
        traffic = None
        try:
            ent = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(top)
            if ent and name == "large":
                traffic = ent["bytes_per_launch"]
        except Exception:
            pass
        roofline = {"bound": bound, "kernel_class": top, "achieved": ach, "peak": peak, "unit": unit,
                    "frac": ach / peak, "traffic": traffic, "peak_source": peaks["src"],
                    "ms_per_step_in_class": ms_step, "launch_groups_per_step": per_class[top][1],
                    "class_ms_per_step": {n: round(v[0], 4) for n, v in per_class.items()},
                    "whole_step_tflops": flops_per_token(c) * T * B / (dev_ms / K * 1e-3) / 1e12}
    line = {
        "metric": METRIC, "value": tokens / (dev_ms * 1e-3), "unit": "tokens/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands, f32 accumulate/state" if args.engine == "tc" else "f32", "data": "synthetic",
        "config": {"workload": workload_name(name, c), "engine": args.engine, "parallelism": f"dp{world}", "dp_transport": getattr(tr, "transport", None),
                   "global_batch": B * world, "seq_len": T,
                   "l2": "no flush: per-step working set (fp32 params+grads 528 MB + activations) exceeds the 126 MB L2"
                   if name == "large" else "no flush; working set may fit L2 for this config"},
        "e2e": {"value": tokens / (e2e_wall_ms * 1e-3), "unit": "tokens/s", "ms_per_step": e2e_wall_ms / K,
                "h2d_bytes_per_step": 2 * T * B * 8, "d2h_bytes_per_step": 8,
                "api": "zaremba_b200.Trainer.train_step_host -> zrb_train_step_host"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "final_loss": loss_after,
        "flops_per_token": flops_per_token(c),
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_port_leg(c, args.cpu_budget)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--config", default="large", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default=os.environ.get("ZRB_ENGINE", "tc"), choices=["tc", "simt"])
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    c = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, c, args.config)
    else:
        run_ours(args, c, args.config)


if __name__ == "__main__":
    main()
