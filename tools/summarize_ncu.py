#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep and the launch-list csv into small text summaries under profiles/."""
import collections, csv, io, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "sm__inst_executed.sum"]

def rep(path, out, tag):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# {tag}: ncu --set full --clock-control none, source {os.path.basename(path)}\n")
        for r in rows[2:]:
            f.write(f"\nkernel: {r[idx['Kernel Name']]}\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"  {k:75s} {r[idx[k]]:>16s} {units[idx[k]]}\n")

def launches(path, out, tag):
    lines = [l for l in open(path).read().split("\n") if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        try:
            k = r["Kernel Name"].split("(")[0][:80]
            a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"])
        except Exception:
            pass
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# {tag}: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n")
        f.write(f"# {len(rows)} launches, {tot/1e6:.3f} ms total\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1]/1e3:10.1f} us {100*v[1]/tot:5.1f}%  n={v[0]:4d}  {k}\n")

if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    go = os.path.join(ROOT, "gpurun_out")
    for f in sorted(os.listdir(go)):
        if f.endswith(".ncu-rep"):
            rep(os.path.join(go, f), os.path.join(ROOT, "profiles", f"{tag}_ncu_{f[5:-8]}.txt"), tag)
    for f in sorted(os.listdir(go)):
        if f.startswith("launches_") and f.endswith(".csv"):
            launches(os.path.join(go, f), os.path.join(ROOT, "profiles", f"{tag}_{f[:-4]}.txt"), tag)
