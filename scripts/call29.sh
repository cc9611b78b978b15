#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shapes_r2.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
BURST_B200_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t1.json 2> gpurun_out/bench_t1.err; grep "compact call" gpurun_out/bench_t1.err | tail -3; python -c "
import json; d=json.load(open('gpurun_out/bench_t1.json')); print('e2e ms:', d['e2e']['ms_per_step'], d['e2e']['value'])"
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t2.json 2> gpurun_out/bench_t2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_t2.json')); print('e2e ms (no timing env):', d['e2e']['ms_per_step'], d['e2e']['value'])"
