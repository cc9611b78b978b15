#!/bin/bash
# ncu --set full on k_filter for one variant (reduced workload)
mkdir -p gpurun_out
V=${1:-1}
BURST_FILTER_VARIANT=$V ncu --set full --clock-control none --import-source on -k regex:k_filter -s 1 -c 1 -o gpurun_out/prof_filter_v$V -f \
    python bench.py --reads 250000 --db-mb 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_filter_v$V.log 2>&1
tail -2 gpurun_out/ncu_filter_v$V.log | cut -c1-200
