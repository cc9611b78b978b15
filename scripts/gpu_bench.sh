#!/bin/bash
# Runs on the GPU box: bench + ncu launch list + one full capture of the top kernels.
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_filter -s 1 -c 1 -o gpurun_out/prof_filter -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_filter.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 5 -c 5 -o gpurun_out/prof_extend -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_extend.log 2>&1
ls -la gpurun_out
