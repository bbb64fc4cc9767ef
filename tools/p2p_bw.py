#!/usr/bin/env python
"""Copy-engine peer-copy bandwidth (GPU1 -> GPU0, pulled by GPU0) vs number of concurrent streams."""
import torch, json
torch.cuda.set_device(0)
n = 33 * 1024 * 1024          # 132 MB of fp32
src = torch.ones(n, device="cuda:1"); dst = torch.empty(n, device="cuda:0")
print("asyncEngineCount", torch.cuda.get_device_properties(0).async_engine_count if hasattr(torch.cuda.get_device_properties(0), "async_engine_count") else "n/a")
out = {}
for k in (1, 2, 4, 8, 16):
    streams = [torch.cuda.Stream(device=0) for _ in range(k)]
    piece = n // k
    def run():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dst[i * piece:(i + 1) * piece].copy_(src[i * piece:(i + 1) * piece], non_blocking=True)
    for _ in range(3): run()
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.default_stream(0))
    for s in streams: s.wait_stream(torch.cuda.default_stream(0))
    for _ in range(5): run()
    for s in streams: torch.cuda.default_stream(0).wait_stream(s)
    e1.record(torch.cuda.default_stream(0)); torch.cuda.synchronize(0)
    ms = e0.elapsed_time(e1) / 5
    out[k] = round(n * 4 / ms / 1e6, 1)
print(json.dumps({"p2p_pull_GBps_by_streams": out}))
