#!/bin/bash
mkdir -p gpurun_out
make -s -C oracle ref >/dev/null 2>&1
timeout 2700 python scripts/manuscript_fixture.py --every 3 --ref-timeout 1500 > gpurun_out/manuscript_fixture.json 2> gpurun_out/manuscript_fixture.err; echo "rc=$?"; tail -c 3000 gpurun_out/manuscript_fixture.json; tail -5 gpurun_out/manuscript_fixture.err
