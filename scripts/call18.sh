#!/bin/bash
mkdir -p gpurun_out
BURST_B200_DEBUG=1 timeout 1200 python scripts/gpu_tune2.py --settings 1:8:0:16:1,1:8:14:16:2:256,1:8:14:16:3:256,1:8:13:16:3:256,1:8:13:16:2:256,1:8:14:16:3:512,1:8:15:16:2,1:8:15:16:3,1:8:14:8:2:256,1:8:14:32:2:256,1:8:14:16:3:128 > gpurun_out/tune3.txt 2> gpurun_out/tune3.err
cat gpurun_out/tune3.txt; grep k_seedw gpurun_out/tune3.err | sort | uniq -c
