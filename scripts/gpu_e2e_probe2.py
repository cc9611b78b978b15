#!/usr/bin/env python
"""Timeline of the compact one-call path (bg_align_bunches_into) on the bench workload (BURST_B200_TIMING=1)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BURST_B200_TIMING"] = "1"
import torch
from burst_b200 import synth
from burst_b200.engine import Engine, MODE_MIN, HIT_DTYPE
w = synth.bunch_workload(1_000_000, 100, 2, 2048 << 20, 214, seed=20261017)
eng = Engine(0); eng.load_db(w["packed"], w["clump_len"])
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory()
    return t.numpy().view(a.dtype).reshape(a.shape), t
budget_r = np.zeros(w["n_reads"], np.uint16); budget_r[w["slot"]] = w["budget"]
c_reads, k1 = pin(Engine.pack2(w["rcodes"])); c_len, k2 = pin(w["rlen"]); c_bud, k3 = pin(budget_r); c_strand, k4 = pin(w["strand"])
c_coff, k5 = pin(w["cand_off"].astype(np.uint32)); c_cand, k6 = pin(w["cand"].astype(np.uint32))
p_hits, k7 = pin(np.zeros(2_100_000, HIT_DTYPE)); p_best, k8 = pin(np.full(w["n_reads"], 0xFFFF, np.uint16))
for it in range(5):
    p_best[:] = 0xFFFF
    t0 = time.perf_counter()
    n = eng.align_bunches_into(c_reads, c_len, c_bud, c_strand, w["qbunch"], c_coff, c_cand, p_hits, p_best, MODE_MIN, packed2=True)
    print("call %d: %.3f ms wall, %d hits" % (it, (time.perf_counter() - t0) * 1e3, n), flush=True)
