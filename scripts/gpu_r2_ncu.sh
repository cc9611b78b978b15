#!/bin/bash
# ncu --set full of the seed kernel and of the first band class of the extend sweep, on a reduced bench workload; the reports are
# summarised on the box (they are too large to bring back) into gpurun_out/*.txt
mkdir -p gpurun_out
# the seed kernel at the FULL bench workload (roofline.traffic is quoted per launch of that workload)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_seedw' -s 1 -c 1 -o /tmp/r2_seed -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r2_seed.log 2>&1
python scripts/ncu_summary.py full /tmp/r2_seed.ncu-rep > gpurun_out/r2_ncu_seed.txt 2>&1
python scripts/ncu_lines.py /tmp/r2_seed.ncu-rep 1000000 0.003 > gpurun_out/r2_lines_seed.txt 2>&1
ncu -i /tmp/r2_seed.ncu-rep --page raw --csv > gpurun_out/r2_raw_seed.csv 2>/dev/null
python scripts/ncu_sass_top.py /tmp/r2_seed.ncu-rep 14 > gpurun_out/r2_sass_seed.txt 2>&1
head -12 gpurun_out/r2_ncu_seed.txt
if [ -z "$SEED_ONLY" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend' -s ${EXT_SKIP:-9} -c 1 -o /tmp/r2_extend -f \
    python bench.py --reads 250000 --db-mb 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r2_extend.log 2>&1
python scripts/ncu_summary.py full /tmp/r2_extend.ncu-rep > gpurun_out/r2_ncu_extend.txt 2>&1
python scripts/ncu_lines.py /tmp/r2_extend.ncu-rep 250000 0.003 > gpurun_out/r2_lines_extend.txt 2>&1
ncu -i /tmp/r2_extend.ncu-rep --page raw --csv > gpurun_out/r2_raw_extend.csv 2>/dev/null
python scripts/ncu_sass_top.py /tmp/r2_extend.ncu-rep 14 > gpurun_out/r2_sass_extend.txt 2>&1
head -12 gpurun_out/r2_ncu_extend.txt
fi
if [ -n "$LAUNCHES" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2_launches.csv 2>&1 | head -40
fi
