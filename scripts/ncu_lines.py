#!/usr/bin/env python
"""Per-source-line instruction counts / stall samples of one kernel from an .ncu-rep (ncu --import-source on, -lineinfo).
usage: ncu_lines.py <file.ncu-rep> [units] [min_share]   -- units = number by which counts are divided (e.g. runs)"""
import csv, subprocess, sys
rep = sys.argv[1]; units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0; mins = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
ia = hdr.index("Instructions Executed"); iss = hdr.index("# Samples")
lines = [(r[0], r[1], int(r[ia]) if r[ia].isdigit() else 0, int(r[iss]) if r[iss].isdigit() else 0) for r in rows if r and r[0].isdigit()]
tot = sum(l[2] for l in lines); ts = sum(l[3] for l in lines)
print("# total warp instructions %d (%.1f per unit), samples %d" % (tot, tot / units, ts))
for ln, src, n, sm in lines:
    if n >= tot * mins or sm >= ts * mins * 2:
        print("%5s %9.1f inst/unit %5.1f%%  samples %5.1f%%  %s" % (ln, n / units, 100.0 * n / tot, 100.0 * sm / max(ts, 1), src.strip()[:110]))
