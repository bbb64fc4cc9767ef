#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/micro/mma_bench > gpurun_out/mma_bench.csv 2>&1; cat gpurun_out/mma_bench.csv
for a in 0 1; do for b in 0 1; do timeout 120 python tools/test_gemm_tc.py $a $b > gpurun_out/gemm_tc_${a}${b}.json 2> gpurun_out/gemm_tc_${a}${b}.err; tail -1 gpurun_out/gemm_tc_${a}${b}.json; done; done
python - <<'PY'
import json
for ab in ("00","01","11"):
    d=json.loads(open(f"gpurun_out/gemm_tc_{ab}.json").readline())
    for c in d["cases"]:
        if c["shape"][0]>=700 or c["shape"][1]>=6000: print(ab, c["shape"], c.get("us"), c.get("tflops"), c.get("ok"))
PY
ZRB_TEST_ENGINES=tc timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --engine tc --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tc_v6.json 2> gpurun_out/bench_tc_v6.err; cat gpurun_out/bench_tc_v6.json; tail -3 gpurun_out/bench_tc_v6.err
