#!/usr/bin/env python
"""Run under torchrun on N GPUs: the CUDA engine behind QuerySharded / ReferenceSharded (NCCL) must
reproduce the single-GPU result bit for bit.  Prints one line per check; exit code 1 on mismatch.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/gpu_multi_check.py"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from burst_b200 import synth, sharded          # noqa: E402
from burst_b200.engine import Engine, RUN_DTYPE  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = synth.bunch_workload(40000, 100, 2, 64 << 20, 214, seed=11)
    runs = np.ascontiguousarray(w["runs"], RUN_DTYPE)
    args = (w["qcodes"], w["qoff"], w["budget"], runs)
    kw = dict(slot=w["slot"], nslots=w["nslots"])
    eng = Engine(local, stream=torch.cuda.current_stream().cuda_stream)
    eng.load_db(w["packed"], w["clump_len"])
    ok = True
    for mode in (0, 1):
        want_hits, want_best = eng.align(w["qcodes"], w["qoff"], w["budget"], None, mode, runs=runs, **kw)
        for kind in ("queries", "references"):
            e2 = Engine(local, stream=torch.cuda.current_stream().cuda_stream)
            drv = (sharded.QuerySharded if kind == "queries" else sharded.ReferenceSharded)(e2)
            drv.load_db(w["packed"], w["clump_len"])
            hits, best = drv.align_runs(*args, mode, **kw)
            good = np.array_equal(hits, want_hits) and np.array_equal(best, want_best)
            ok &= good
            if rank == 0:
                print("multi-gpu check: world=%d kind=%s mode=%d hits=%d %s" % (world, kind, mode, len(hits), "OK" if good else "MISMATCH"), flush=True)
            e2.close()
    t = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(t)
    dist.destroy_process_group()
    sys.exit(1 if int(t.item()) else 0)


if __name__ == "__main__":
    main()
