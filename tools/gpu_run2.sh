#!/bin/bash
mkdir -p gpurun_out
for a in 0 1; do for b in 0 1; do
  timeout 120 python tools/test_gemm_tc.py $a $b > gpurun_out/gemm_tc_${a}${b}.json 2> gpurun_out/gemm_tc_${a}${b}.err; echo "gemm $a $b rc=$?"
  tail -1 gpurun_out/gemm_tc_${a}${b}.json; tail -2 gpurun_out/gemm_tc_${a}${b}.err
done; done
ZRB_TEST_ENGINES=simt timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
