#!/bin/bash
# round-2 closing run on one GPU: GPU tier, smoke, the default bench line (CPU baseline + parity), ncu --set full of k_seedw on the full
# bench workload with per-line / SASS summaries, launch list of a bench run
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench.json")); e = d["e2e"]
print("value %.0f M ms %.3f | e2e %.0f M %.3f ms (%d in flight) | one call %.0f M %.3f ms | frac %.3f filter %.3f extend %.3f | cpu %s | parity %s" % (
    d["value"] / 1e6, d["ms_per_step"], e["value"] / 1e6, e["ms_per_step"], e["contexts_in_flight"], e["one_call_at_a_time"]["value"] / 1e6, e["one_call_at_a_time"]["ms_per_step"],
    d["roofline"]["frac"], d["work"]["ms_filter"], d["work"]["ms_extend"], d.get("cpu_baseline", {}).get("value"), d.get("parity")))
PY
SEED_ONLY=1 LAUNCHES=1 bash scripts/gpu_r2_ncu.sh 2>&1 | tail -45
