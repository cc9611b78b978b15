#!/bin/bash
# Round 2 iteration call: GPU tests (stop at first failure), the seed-filter sweep, optionally a bench line and an ncu capture.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python scripts/gpu_tune2.py --settings "${SETTINGS:-0:8:0:8:1,1:8:0:8:1,1:8:0:8:2,1:4:0:8:1,1:4:0:8:2,1:8:16:8:1,1:8:14:8:2,1:8:0:16:1}" > gpurun_out/tune2.txt 2>&1; cat gpurun_out/tune2.txt
if [ -n "$BENCH" ]; then timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['work']['ms_extend'], d.get('parity'))
PY
tail -3 gpurun_out/bench_iter.err; fi
if [ -n "$NCU" ]; then bash scripts/gpu_r2_ncu.sh; fi
