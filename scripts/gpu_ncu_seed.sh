#!/bin/bash
# ncu --set full on k_seed (reduced workload so the ~40 replays stay short) + launch list of a bench run
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_seed -s 1 -c 1 -o gpurun_out/prof_seed -f \
    python bench.py --reads 250000 --db-mb 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_seed.log 2>&1
tail -2 gpurun_out/ncu_seed.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
