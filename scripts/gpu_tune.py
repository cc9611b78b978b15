#!/usr/bin/env python
"""Kernel-only timings of the bench workload under different tuning knobs (runs on the GPU box).
usage: python scripts/gpu_tune.py [--reads N] [--db-mb M]  -> one line per setting: ms_filter / ms_extend / ms_select"""
import argparse, os, sys, itertools
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from burst_b200 import synth
from burst_b200.engine import Engine, MODE_MIN, RUN_DTYPE, PARAM_SEED_CHUNK, PARAM_SEED_WORDS, PARAM_SEED_STAGE, PARAM_SEED_GROUPS

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=1_000_000)
ap.add_argument("--db-mb", type=int, default=2048)
ap.add_argument("--settings", default="8:0:1,8:1024:1,8:512:1,8:1024:0,16:1024:1,4:1024:1,32:1024:1")
a = ap.parse_args()
w = synth.bunch_workload(a.reads, 100, 2, a.db_mb << 20, 214, seed=20261017)
eng = Engine(0)
eng.load_db(w["packed"], w["clump_len"])
runs = np.ascontiguousarray(w["runs"], RUN_DTYPE)
for s in a.settings.split(","):
    f = [int(x) for x in s.split(":")]
    chunk, words, stage = f[:3]; groups = f[3] if len(f) > 3 else 0
    eng.set_param(PARAM_SEED_CHUNK, chunk); eng.set_param(PARAM_SEED_WORDS, words); eng.set_param(PARAM_SEED_STAGE, stage); eng.set_param(PARAM_SEED_GROUPS, groups)
    eng.upload_runs(w["qcodes"], w["qoff"], w["budget"], runs, slot=w["slot"], nslots=w["nslots"])
    best = None
    for it in range(4):
        eng.run(MODE_MIN); eng.count(); st = eng.stats()
        if best is None or st["ms_filter"] < best["ms_filter"]:
            best = st
    print("groups %d chunk %3d words %5d stage %d : filter %.3f ms  extend %.3f ms  select %.3f ms  survivors %d hits %d" % (
        groups, chunk, best["seed_words"], stage, best["ms_filter"], best["ms_extend"], best["ms_select"], best["survivors"], best["hits"]), flush=True)
