#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes_r2.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 1500 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c3.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['work'])
PY
tail -3 gpurun_out/bench_c3.err
timeout 600 python scripts/gpu_tune2.py --settings 1:8:0:8:2,1:8:0:8:2 2>&1 | cut -c40-140
