#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/manuscript_fixture.py --every 3 --skip-reference > /dev/null 2>&1
cd /tmp/ms
run() { echo "== $1"; env $1 BURST_B200_DEBUG=1 BURST_B200_TIMING=1 /root/repo/burst_b200/host/burst-b200 -r ms.edx -a ms.acx -q genes.fna -m ALLPATHS -i 0.98 --noprogress -o x.b6 -t 16 2>&1 | grep -E "extend:|one-call timing|search  " | sed 's/.*c7\[[^]]*\]//; s/.*one-call timing: //' | tail -9; sort x.b6 | md5sum; }
run "BURST_B200_GEN_MODE=1"
run "BURST_B200_GEN_MODE=0"
