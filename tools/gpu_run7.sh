#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log; tail -12 gpurun_out/pytest_all.log
timeout 120 python tools/rec_trace.py large > gpurun_out/rec_trace_large.json 2> gpurun_out/rec_trace.err; cat gpurun_out/rec_trace_large.json; tail -2 gpurun_out/rec_trace.err
timeout 300 python bench.py --engine tc --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tc_v5.json 2> gpurun_out/bench_tc_v5.err; cat gpurun_out/bench_tc_v5.json; tail -3 gpurun_out/bench_tc_v5.err
