#!/bin/bash
# Runs on the GPU box: GPU parity tests, smoke, then a bench run (no CPU baseline unless asked).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS:---no-cpu-baseline} > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
