#!/bin/bash
# after a change to the one-call preparation: GPU tier, e2e timeline, bench with 3 contexts in flight
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python scripts/gpu_e2e_probe2.py 2>&1 | tail -4
for n in 3 2; do
timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --inflight $n > gpurun_out/bench_inflight$n.json 2> gpurun_out/bench_inflight$n.err; echo "inflight $n rc=$?"; tail -2 gpurun_out/bench_inflight$n.err
python -c "
import json;d=json.load(open('gpurun_out/bench_inflight$n.json'));e=d['e2e'];print('value',round(d['value']/1e6),'ms',round(d['ms_per_step'],3),'e2e',round(e['value']/1e6),round(e['ms_per_step'],3),'one',round(e['one_call_at_a_time']['value']/1e6),round(e['one_call_at_a_time']['ms_per_step'],3))"
done
