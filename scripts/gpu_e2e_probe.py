#!/usr/bin/env python
"""Where the end-to-end time goes: one-call path under different slice counts, resident API per call, PCIe copy rates."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from burst_b200 import synth
from burst_b200.engine import Engine, MODE_MIN, RUN_DTYPE, HIT_DTYPE, PARAM_PIPE_SLICES, PARAM_PIPE_RATIO
w = synth.bunch_workload(1_000_000, 100, 2, 2048 << 20, 214, seed=20261017)
eng = Engine(0); eng.load_db(w["packed"], w["clump_len"])
runs = np.ascontiguousarray(w["runs"], RUN_DTYPE)
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory()
    return t.numpy().view(a.dtype).reshape(a.shape), t
pc, k1 = pin(w["qcodes"]); po, k2 = pin(w["qoff"]); pb, k3 = pin(w["budget"]); ps, k4 = pin(w["slot"]); pr, k5 = pin(runs)
ph, k6 = pin(np.zeros(2_000_000, HIT_DTYPE)); pbest, k7 = pin(np.full(w["nslots"], 0xFFFF, np.uint16))
for mb in (64, 256):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory(); d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print("pinned copy %4d MB: H2D %.1f GB/s" % (mb, mb / 1024 / dt), flush=True)
for it in range(3):
    t0 = time.perf_counter(); eng.upload_runs(pc, po, pb, pr, slot=ps, nslots=w["nslots"]); t1 = time.perf_counter()
    eng.run(MODE_MIN); n = eng.count(); t2 = time.perf_counter()
    hits, best = eng.download(); t3 = time.perf_counter()
    print("resident: upload %.2f ms  run+count %.2f ms  download %.2f ms  (hits %d)" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, n), flush=True)
PACK = ("packed4", pin(Engine.pack4(w["qcodes"]))[0]) if os.environ.get("PACKED") else None
_keep = PACK
if PACK is not None:
    pp, kk = pin(Engine.pack4(w["qcodes"])); PACK = ("packed4", pp)
    pc = PACK
for slices, ratio in ((4, 140), (4, 140), (4, 100), (4, 120), (5, 110), (3, 120), (6, 100), (8, 100), (5, 130)):
    eng.set_param(PARAM_PIPE_SLICES, slices); eng.set_param(PARAM_PIPE_RATIO, ratio)
    ts = []
    for it in range(4):
        pbest[:] = 0xFFFF
        t0 = time.perf_counter(); n = eng.align_runs_into(pc, po, pb, pr, ph, pbest, MODE_MIN, slot=ps, nslots=w["nslots"]); ts.append((time.perf_counter() - t0) * 1e3)
    print("one call, %2d slices ratio %d: %s ms (hits %d)" % (slices, ratio, " ".join("%.2f" % t for t in ts), n), flush=True)
    t0 = time.perf_counter(); n = eng.align_runs_into(pc, po, pb, pr, ph, None, MODE_MIN, slot=ps, nslots=w["nslots"]); print("   without best in/out: %.2f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
