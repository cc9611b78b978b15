#!/bin/bash
# run the tune workload once per library variant under burst_b200/variants/
for v in ${VARIANTS:-v0 v1 v2 v3 v4 v6}; do
  cp burst_b200/variants/$v.so burst_b200/libburst_b200.so
  echo "== $v: $(python scripts/gpu_tune.py --settings ${SETTINGS:-8:0:0} 2>&1 | tail -1 | cut -c1-150)"
done
