#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench8 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_8gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['config'].get('numa_node_rank0'))
PY
tail -3 gpurun_out/bench_8gpu.err; nproc; free -g | head -2
