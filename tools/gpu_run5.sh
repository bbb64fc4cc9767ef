#!/bin/bash
mkdir -p gpurun_out
ZRB_TEST_ENGINES=tc timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tc.log
tail -25 gpurun_out/pytest_tc.log
timeout 300 python bench.py --engine tc --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tc_v3.json 2> gpurun_out/bench_tc_v3.err; cat gpurun_out/bench_tc_v3.json; tail -3 gpurun_out/bench_tc_v3.err
