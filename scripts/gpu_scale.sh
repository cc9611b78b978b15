#!/bin/bash
# bench.py at N GPUs (NGPU, default 8) under torch.distributed.run, as the driver launches it
N=${NGPU:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "rc=$? bytes=$(wc -c < gpurun_out/r2_bench_${N}gpu.json)"
tail -5 gpurun_out/r2_bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_${N}gpu.json"))
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "packed4", d["e2e"]["packed4_strands"]["value"])
except Exception as e:
    print("no json:", e)
PY
