#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log
tail -15 gpurun_out/pytest_all.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --engine tc --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tc_v1.json 2> gpurun_out/bench_tc_v1.err; cat gpurun_out/bench_tc_v1.json; tail -3 gpurun_out/bench_tc_v1.err
