#!/bin/bash
# Runs on the GPU box: time k_filter variants on a reduced configs[1] workload.
mkdir -p gpurun_out
for v in ${VARIANTS:-1 2 3}; do
  BURST_FILTER_VARIANT=$v python bench.py --reads 250000 --db-mb 512 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/variant_$v.json 2> gpurun_out/variant_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/variant_$v.json"))
print("variant $v: ms_filter=%.3f ms_extend=%.3f value=%.0f e2e=%.0f found=%d planted=%d" % (d["work"]["ms_filter"], d["work"]["ms_extend"], d["value"], d["e2e"]["value"], d["work"]["reads_found"], d["work"]["reads_at_planted_lane"]))
PY
done
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
