#!/bin/bash
# GPU call 16: device candidate generation -- parity tests, CLI golden with the flag, whole-binary with/without the flag
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_acx_search.py tests/test_gpu_golden.py -x -q -m gpu > gpurun_out/pytest_acx.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_acx.log
make -s -C oracle ref >/dev/null 2>&1
timeout 900 python scripts/whole_binary.py --shape shotgun --mbp 100 --reads 1000000 --skip-reference > gpurun_out/wb_shotgun_host.json 2> gpurun_out/wb_err.log; tail -c 1500 gpurun_out/wb_shotgun_host.json
cp /tmp/wb/shotgun/ours.b6 /tmp/wb/shotgun/host.b6
timeout 900 python scripts/whole_binary.py --shape shotgun --mbp 100 --reads 1000000 --skip-reference --ours-extra=--device-candidates > gpurun_out/wb_shotgun_dev.json 2>> gpurun_out/wb_err.log; tail -c 1500 gpurun_out/wb_shotgun_dev.json
sort /tmp/wb/shotgun/ours.b6 | md5sum; sort /tmp/wb/shotgun/host.b6 | md5sum
tail -5 gpurun_out/wb_err.log
