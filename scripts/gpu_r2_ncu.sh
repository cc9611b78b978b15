#!/bin/bash
# ncu --set full of the seed kernel and of the first band class of the extend sweep, on a reduced bench workload
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_seed' -s 1 -c 1 -o gpurun_out/r2_seed -f \
    python bench.py --reads 250000 --db-mb 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r2_seed.log 2>&1
tail -2 gpurun_out/ncu_r2_seed.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend' -s ${EXT_SKIP:-9} -c 1 -o gpurun_out/r2_extend -f \
    python bench.py --reads 250000 --db-mb 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r2_extend.log 2>&1
tail -2 gpurun_out/ncu_r2_extend.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2_launches.csv 2>&1 | head -40
