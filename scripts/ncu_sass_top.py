#!/usr/bin/env python
"""Hottest SASS instructions (by stall samples) of one kernel from an .ncu-rep, each with the instructions just before it.
usage: ncu_sass_top.py <file.ncu-rep> [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and "Source" in r and "# Samples" in r)
isrc = hdr.index("Source"); ismp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
body = [r for r in rows[rows.index(hdr) + 1:] if len(r) > max(isrc, ismp)]
def num(x):
    try: return int(x)
    except ValueError: return 0
tot = sum(num(r[ismp]) for r in body) or 1
top = sorted(range(len(body)), key=lambda i: -num(body[i][ismp]))[:n]
for i in sorted(top):
    print("---- %.1f%% of samples" % (100.0 * num(body[i][ismp]) / tot))
    for j in range(max(0, i - 6), i + 1):
        print("   %s %6d smp %9d exe  %s" % (">>" if j == i else "  ", num(body[j][ismp]), num(body[j][iex]), body[j][isrc][:110]))
