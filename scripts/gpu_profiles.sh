#!/bin/bash
# Runs on the GPU box: everything profiles/ is summarised from (bench line, launch list, full captures of the two top
# kernels on a reduced workload so the ~40 replays stay short, DRAM traffic of one launch of each on the bench workload).
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 600 gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_seed|k_extend' -s 6 -c 6 --csv \
    --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/traffic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_seed -s 1 -c 1 -o gpurun_out/prof_seed -f \
    python bench.py --reads 250000 --db-mb 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_prof_seed.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 5 -c 1 -o gpurun_out/prof_extend -f \
    python bench.py --reads 250000 --db-mb 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_prof_extend.log 2>&1
ls -la gpurun_out | tail -12
