#!/bin/bash
mkdir -p gpurun_out
BURST_B200_TIMING=1 timeout 1500 python scripts/manuscript_fixture.py --every 3 --skip-reference > gpurun_out/manuscript_timing.json 2> gpurun_out/manuscript_timing.err; echo "rc=$?"; tail -c 2500 gpurun_out/manuscript_timing.json; tail -5 gpurun_out/manuscript_timing.err
