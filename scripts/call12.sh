mkdir -p gpurun_out
BURST_B200_EXT_STAGE=1 python scripts/gpu_tune2.py --settings 1:8:0:16:1 2>&1 | tail -1
BURST_B200_EXT_STAGE=0 python scripts/gpu_tune2.py --settings 1:8:0:16:1 2>&1 | tail -1
SEED_ONLY= bash scripts/gpu_r2_ncu.sh > /dev/null 2>&1
cat gpurun_out/r2_sass_extend.txt | cut -c1-150
