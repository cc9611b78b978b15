#!/bin/bash
mkdir -p gpurun_out
make -s -C oracle ref >/dev/null 2>&1
timeout 900 python scripts/whole_binary.py --shape amplicon --mbp 28 --reads 200000 --ours-extra=--device-candidates > gpurun_out/wb_amplicon_dev.json 2> gpurun_out/wb_err.log; tail -c 1800 gpurun_out/wb_amplicon_dev.json; tail -3 gpurun_out/wb_err.log
timeout 1700 python bench.py --config target --steps 5 --warmup 3 > gpurun_out/bench_target.json 2> gpurun_out/bench_target.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_target.json; tail -5 gpurun_out/bench_target.err
