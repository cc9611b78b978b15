#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['work']['ms_extend'], d.get('parity'), d['clocks'], d['cpu_baseline']['value'])
PY
make -s -C oracle ref >/dev/null 2>&1
timeout 900 python scripts/whole_binary.py --shape shotgun --mbp 100 --reads 1000000 --ours-extra=--device-candidates > gpurun_out/wb_shotgun_dev.json 2> gpurun_out/wb_err.log; tail -c 1500 gpurun_out/wb_shotgun_dev.json; echo
