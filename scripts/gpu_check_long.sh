#!/bin/bash
# long-query shapes after a change to k_filter / k_extend<0>: the GPU tier, then the reference's manuscript data set through both binaries
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
make -s -C oracle ref >/dev/null 2>&1
BURST_B200_TIMING=1 timeout 1500 python scripts/manuscript_fixture.py --every ${EVERY:-3} --ref-timeout 600 > gpurun_out/manuscript_fixture.json 2> gpurun_out/manuscript_fixture.err; echo "rc=$?"; tail -c 2000 gpurun_out/manuscript_fixture.json; tail -5 gpurun_out/manuscript_fixture.err
