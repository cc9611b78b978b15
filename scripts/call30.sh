#!/bin/bash
mkdir -p gpurun_out
NGPU=2 bash scripts/gpu_multi_cli.sh > gpurun_out/multi_cli_2gpu.txt 2>&1; echo "multi cli rc=$?"; cat gpurun_out/multi_cli_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e']['ms_per_step'])
PY
