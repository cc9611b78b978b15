#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/manuscript_fixture.py --every 12 --skip-reference > gpurun_out/ms12.json 2>/dev/null; tail -c 900 gpurun_out/ms12.json; echo
cd /tmp/ms
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file /root/repo/gpurun_out/ms_launches.csv /root/repo/burst_b200/host/burst-b200 -r ms.edx -a ms.acx -q genes.fna -m ALLPATHS -i 0.98 --noprogress -o ncu.b6 -t 16 > /root/repo/gpurun_out/ms_ncu.log 2>&1
cd /root/repo
python scripts/ncu_summary.py launches gpurun_out/ms_launches.csv 2>&1 | head -50
rm -f gpurun_out/ms_launches.csv
