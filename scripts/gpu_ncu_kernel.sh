#!/bin/bash
# ncu --set full on one kernel (KREGEX, default k_seed) of a reduced bench workload (so the ~40 replays stay short)
K=${KREGEX:-k_seed}; OUT=${OUT:-prof_seed}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-1} -c 1 -o gpurun_out/$OUT -f \
    python bench.py --reads 250000 --db-mb 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$OUT.log 2>&1
tail -2 gpurun_out/ncu_$OUT.log | cut -c1-300
