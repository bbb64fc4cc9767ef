#!/bin/bash
mkdir -p gpurun_out
for ab in "0 0" "0 1" "1 1"; do set -- $ab; timeout 120 python tools/test_gemm_tc.py $1 $2 > gpurun_out/gemm_tc_$1$2.json 2> gpurun_out/gemm_tc_$1$2.err; tail -1 gpurun_out/gemm_tc_$1$2.json; done
python - <<'PY'
import json
for ab in ("00","01","11"):
    d=json.loads(open(f"gpurun_out/gemm_tc_{ab}.json").readline())
    for c in d["cases"]:
        if c["shape"][0]>=700 or c["shape"][1]>=6000: print(ab, c["shape"], c.get("us"), c.get("tflops"), c.get("ok"))
PY
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tc_v7.json 2> gpurun_out/bench_tc_v7.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tc_v7.json')); print(d['ms_per_step'], d['value'], d['roofline']['class_ms_per_step'])"; tail -2 gpurun_out/bench_tc_v7.err
