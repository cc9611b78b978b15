#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --config c3 --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"; tail -c 3500 gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "c2 rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['work']['ms_extend'], d.get('parity'))
PY
