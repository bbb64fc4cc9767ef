#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/micro/mma_bench > gpurun_out/mma_bench.csv 2>&1; cat gpurun_out/mma_bench.csv
