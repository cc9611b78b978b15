#!/usr/bin/env python
"""Which golden kernel vectors fail, under which engine settings (debugging aid, runs on the GPU box)."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from burst_b200.engine import Engine, default_scoring, MODE_MIN
    z = int(sys.argv[2])
    g = np.load(os.path.join(ROOT, "tests", "golden", "kernels_z%d.npz" % z))
    eng = Engine(0); eng.set_scoring(default_scoring(z))
    bad = 0
    only = [int(x) for x in os.environ.get("ONLY", "").split(",") if x]
    for i in (only or range(len(g["clen"]))):
        if only: os.environ["BURST_B200_DEBUG"] = "1"
        packed = g["packed"][g["packed_off"][i]:g["packed_off"][i + 1]]
        q = g["q"][g["q_off"][i]:g["q_off"][i + 1]]
        emac, rm = int(g["emac"][i]), int(g["min"][i])
        eng.load_db(packed, np.array([g["clen"][i]], np.uint32))
        hits, best = eng.align(q, np.array([0, len(q)], np.uint64), np.array([emac], np.uint16), None, MODE_MIN)
        want = 0xFFFF if (rm == 0xFFFFFFFF or rm > emac) else rm
        if int(best[0]) != want:
            st = eng.stats(); bad += 1
            if bad <= 4:
                print("  vector %d: qlen %d clen %d emac %d want %d got %d  survivors %d seed_queries %d stride %d window %d band_cells %d" % (
                    i, len(q), int(g["clen"][i]), emac, want, int(best[0]), st["survivors"], st["seed_queries"], st["seed_stride"], st["seed_window"], st["band_cells"]))
    print("  z=%d: %d of %d vectors wrong" % (z, bad, len(g["clen"])))
else:
    for env in ({"ONLY": "2,3"}, {"BURST_B200_EXT_STAGE": "0", "ONLY": "2,3"}):
        print("settings", env, flush=True)
        subprocess.run([sys.executable, __file__, "child", "0"], env=dict(os.environ, **env))
