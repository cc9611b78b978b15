#!/usr/bin/env python
"""Round-2 kernel-only timings of the bench workload under the seed-filter knobs (runs on the GPU box).
usage: python scripts/gpu_tune2.py [--reads N] [--db-mb M] [--settings impl:nch:lbits:chunk:fb[:hslots],...]
One line per setting: ms_filter / ms_extend / ms_select, and whether hits + minima equal the first setting's."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from burst_b200 import synth
from burst_b200.engine import (Engine, MODE_MIN, RUN_DTYPE, PARAM_SEED_CHUNK, PARAM_SEED_IMPL, PARAM_SEED_NCH, PARAM_SEED_LBITS, PARAM_SEED_FB, PARAM_SEED_HSLOTS, PARAM_SEED_VMODE)

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=1_000_000)
ap.add_argument("--db-mb", type=int, default=2048)
ap.add_argument("--settings", default="0:8:0:8:1,1:8:0:8:1,1:8:0:8:2,1:4:0:8:1,1:8:16:8:1,1:8:14:8:2,1:8:0:16:1")
ap.add_argument("--lib", default=None, help="engine library to load instead of burst_b200/libburst_b200.so (A/B builds)")
a = ap.parse_args()
w = synth.bunch_workload(a.reads, 100, 2, a.db_mb << 20, 214, seed=20261017)
eng = Engine(0, lib_path=a.lib)
eng.load_db(w["packed"], w["clump_len"])
runs = np.ascontiguousarray(w["runs"], RUN_DTYPE)
ref = None
for s in a.settings.split(","):
    f = [int(x) for x in s.split(":")]
    impl, nch, lbits, chunk, fb = f[:5]; hs = f[5] if len(f) > 5 else 0; vm = f[6] if len(f) > 6 else 0
    eng.set_param(PARAM_SEED_VMODE, vm)
    eng.set_param(PARAM_SEED_HSLOTS, hs)
    eng.set_param(PARAM_SEED_IMPL, impl); eng.set_param(PARAM_SEED_NCH, nch); eng.set_param(PARAM_SEED_LBITS, lbits); eng.set_param(PARAM_SEED_CHUNK, chunk); eng.set_param(PARAM_SEED_FB, fb)
    eng.upload_runs(w["qcodes"], w["qoff"], w["budget"], runs, slot=w["slot"], nslots=w["nslots"])
    best = None
    for it in range(4):
        eng.run(MODE_MIN); eng.count(); st = eng.stats()
        if best is None or st["ms_filter"] + st["ms_extend"] < best["ms_filter"] + best["ms_extend"]:
            best = st
    hits, mins = eng.download()
    if ref is None:
        ref = (hits, mins); same = "reference"
    else:
        same = "same" if (len(hits) == len(ref[0]) and np.array_equal(hits, ref[0]) and np.array_equal(mins, ref[1])) else "DIFFERENT (%d vs %d hits)" % (len(hits), len(ref[0]))
    print("vm %d " % vm, end=""); print("impl %d nch %d lbits %2d chunk %3d fb %d hs %4d : filter %.3f ms  extend %.3f ms  select %.3f ms  survivors %d hits %d  [%s]" % (
        impl, nch, lbits, chunk, fb, hs, best["ms_filter"], best["ms_extend"], best["ms_select"], best["survivors"], best["hits"], same), flush=True)
