#!/bin/bash
mkdir -p gpurun_out
make -s -C oracle ref >/dev/null 2>&1
timeout 900 python scripts/whole_binary.py --shape shotgun --mbp 100 --reads 1000000 --ours-extra=--device-candidates > gpurun_out/wb_shotgun_dev.json 2> gpurun_out/wb_err.log; tail -c 1700 gpurun_out/wb_shotgun_dev.json; echo
timeout 900 python scripts/whole_binary.py --shape shotgun --mbp 100 --reads 1000000 --reuse --skip-reference > gpurun_out/wb_shotgun_host.json 2>> gpurun_out/wb_err.log; tail -c 1300 gpurun_out/wb_shotgun_host.json; echo
timeout 900 python scripts/whole_binary.py --shape amplicon --mbp 28 --reads 200000 > gpurun_out/wb_amplicon.json 2>> gpurun_out/wb_err.log; tail -c 1500 gpurun_out/wb_amplicon.json; echo
tail -3 gpurun_out/wb_err.log
