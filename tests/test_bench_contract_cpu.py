"""The bench contract that can be checked without a GPU: `bench.py --impl reference` (the reference's CPU path through the
torch port) prints exactly ONE JSON line on stdout with the keys the driver reads, whatever libraries print meanwhile."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "small",
                        "--steps", "3", "--warmup", "3"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, OMP_NUM_THREADS="1"))          # what torchrun would export
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    # the arm pins torch's own default thread count (one per physical core) even when OMP_NUM_THREADS=1 is inherited
    assert d["cpu_baseline"]["cores"] == max(1, len(os.sched_getaffinity(0)) // 2)
    assert set(d["config"]) >= {"workload", "device", "engine", "parallelism", "global_batch", "seq_len"}
