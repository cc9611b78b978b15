#!/bin/bash
# k_seedw item size (chunks of 32 columns per staged item) A/B on the bench workload, then the GPU tier
mkdir -p gpurun_out
timeout 600 python scripts/gpu_tune2.py --settings 1:8:0:8:2,1:7:0:8:2,1:6:0:8:2,1:5:0:8:2,1:4:0:8:2,1:0:0:8:2,1:8:0:8:2,1:0:0:8:2 > gpurun_out/tune_nch.txt 2> gpurun_out/tune_nch.err; echo "tune rc=$?"; cat gpurun_out/tune_nch.txt; tail -3 gpurun_out/tune_nch.err
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
