#!/bin/bash
# Round 2, first GPU call: box facts, the GPU test tier under the new kernels (and, if it fails, under the old seed
# kernel to separate the two changes), memcheck of the smoke case, the seed-filter tuning sweep, one bench line.
mkdir -p gpurun_out
{ nvidia-smi -L; nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA node\(s\)"; nvidia-smi topo -m 2>/dev/null | head -12; } > gpurun_out/box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
if ! grep -q " passed" gpurun_out/pytest_gpu.log || grep -q "failed" gpurun_out/pytest_gpu.log; then
  BURST_B200_SEED_IMPL=0 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_oldseed.log 2>&1; echo "pytest(old seed) rc=$?"
  tail -5 gpurun_out/pytest_gpu_oldseed.log
fi
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke.log 2>&1; tail -3 gpurun_out/memcheck_smoke.log
timeout 900 python scripts/gpu_tune2.py > gpurun_out/tune2.txt 2>&1; cat gpurun_out/tune2.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 1500 gpurun_out/bench_r2a.json; tail -3 gpurun_out/bench_r2a.err
