#!/bin/bash
# the e2e leg with 1, 2, 3 contexts in flight (bg_share_db), then the two-context test
mkdir -p gpurun_out
for n in 2 3 1; do
  timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --inflight $n > gpurun_out/bench_inflight$n.json 2> gpurun_out/bench_inflight$n.err; echo "inflight $n rc=$?"; tail -2 gpurun_out/bench_inflight$n.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_inflight$n.json'));e=d['e2e'];print('value',round(d['value']/1e6),'ms',round(d['ms_per_step'],3),'e2e',round(e['value']/1e6),round(e['ms_per_step'],3),'one',round(e['one_call_at_a_time']['value']/1e6),round(e['one_call_at_a_time']['ms_per_step'],3))"
done
timeout 300 python -m pytest tests/test_gpu_shapes_r2.py -m gpu -x -q -k "two_contexts" 2>&1 | tail -3
