#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/micro/pdl_overlap_bench > gpurun_out/pdl_overlap_micro.txt 2>&1; cat gpurun_out/pdl_overlap_micro.txt
export ZRB_ERROR_REPORT=gpurun_out/r02_error_at_baseline_configs.json
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log; tail -25 gpurun_out/pytest_all.log
