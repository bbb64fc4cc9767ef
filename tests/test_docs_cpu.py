"""Docs stay in step with the code: every ZRB_* environment switch the library, the trainer or bench.py reads is listed in
INTEGRATION.md's table, and every entry point the header declares is mentioned in INTEGRATION.md or DESIGN.md."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*parts):
    with open(os.path.join(ROOT, *parts)) as f:
        return f.read()


def test_every_environment_switch_is_documented():
    srcs = glob.glob(os.path.join(ROOT, "zaremba_b200", "csrc", "*")) + glob.glob(os.path.join(ROOT, "zaremba_b200", "*.py"))
    srcs.append(os.path.join(ROOT, "bench.py"))
    used = set()
    for p in srcs:
        if os.path.isfile(p):
            with open(p, errors="ignore") as f:
                txt = f.read()
            used |= set(re.findall(r'getenv\("(ZRB_[A-Z0-9_]+)"\)', txt))
            used |= set(re.findall(r'environ(?:\.get)?[\(\[]"(ZRB_[A-Z0-9_]+)"', txt))
    doc = _read("INTEGRATION.md")
    missing = sorted(v for v in used if v not in doc)
    assert not missing, f"undocumented environment switches: {missing}"


def test_public_entry_points_are_mentioned_in_the_docs():
    header = _read("include", "zaremba_b200.h")
    names = set(re.findall(r"\b(zrb_[a-z0-9_]+)\s*\(", header))
    docs = _read("INTEGRATION.md") + _read("DESIGN.md") + _read("README.md")
    # accessors / profiling helpers that only the header documents
    header_only = {n for n in names if n.startswith(("zrb_prof_", "zrb_dp_")) or n in {
        "zrb_version", "zrb_launch_count", "zrb_ctx_workspace_bytes", "zrb_last_error"}}
    missing = sorted(n for n in names - header_only if n not in docs)
    assert not missing, f"entry points never mentioned outside the header: {missing}"
