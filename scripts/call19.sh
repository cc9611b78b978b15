#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['work']['ms_extend'], d.get('parity'))
PY
tail -3 gpurun_out/bench_iter.err
LAUNCHES=1 bash scripts/gpu_r2_ncu.sh
