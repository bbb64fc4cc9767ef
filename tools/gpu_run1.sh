#!/bin/bash
# first GPU session: probe, parity tests (simt engine), cuDNN baseline, bench, ncu launch list
mkdir -p gpurun_out
{
echo "== probe"; nproc; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv; ls /root/reference 2>&1 | head -3; python -c "import torch;print(torch.__version__, torch.backends.cudnn.version(), torch.cuda.nccl.version())"; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"
} > gpurun_out/probe.txt 2>&1
ZRB_TEST_ENGINES=simt timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_simt.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_simt.log
timeout 600 python tests/baseline_cudnn.py --steps 30 > gpurun_out/baseline_cudnn.json 2> gpurun_out/baseline_cudnn.err
timeout 600 python bench.py --engine simt --steps 5 --warmup 3 > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_simt.csv python bench.py --engine simt --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/pytest_simt.log; cat gpurun_out/bench_simt.json | head -c 3000; echo; tail -3 gpurun_out/bench_simt.err
