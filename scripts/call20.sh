#!/bin/bash
mkdir -p gpurun_out
{
for cfg in "0 5" "0 4" "0 3" "24 4" "56 4" "56 3" "56 2" "120 3" "120 2"; do set -- $cfg
echo "== pad $1 bps $2"; BURST_B200_STAGE_PAD=$1 BURST_B200_EXT_BPS=$2 timeout 600 python scripts/gpu_tune2.py --settings 1:8:0:8:2 2>&1 | cut -c40-130
done
} > gpurun_out/ab_extend2.txt 2>&1
cat gpurun_out/ab_extend2.txt
