#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show what the shipped library runs on (DPX add-min, 1-D TMA bulk copies and
mbarriers, cp.async staging, shared-memory atomics), from `cuobjdump -sass burst_b200/libburst_b200.so`.  No GPU needed.
usage: python scripts/sass_excerpt.py > profiles/r2_sass_excerpt.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "burst_b200", "libburst_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["VIADDMNMX", "VIMNMX3", "VIMNMX", "UBLKCP", "SYNCS", "LDGSTS", "LDGDEPBAR", "ATOMS", "ATOMG", "REDG", "RED.", "LDS", "STS", "LDG", "STG", "SHFL", "POPC", "IMAD", "LOP3", "SHF", "HMMA", "UTCMMA", "UTMALDG"]
kern = None; counts = collections.OrderedDict(); arch = ""
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts.setdefault(kern, collections.Counter()); continue
    m = re.search(r"arch = (\S+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1); counts[kern]["_all"] += 1
        for w in WANT:
            if op.startswith(w):
                counts[kern][w] += 1; break
print("# cuobjdump -sass burst_b200/libburst_b200.so (%s): instruction counts per kernel (static)" % arch)
print("# VIADDMNMX/VIMNMX3 = DPX add-min / 3-input min; UBLKCP = cp.async.bulk (1-D TMA); SYNCS = mbarrier ops; LDGSTS = cp.async; no tensor-core ops (HMMA/UTCMMA): integer min-plus")
cols = ["_all", "VIADDMNMX", "VIMNMX3", "VIMNMX", "UBLKCP", "SYNCS", "LDGSTS", "ATOMS", "ATOMG", "REDG", "SHFL", "LDS", "LDG", "HMMA", "UTCMMA"]
print("%-44s" % "kernel" + "".join("%10s" % c for c in cols))
tot = collections.Counter()
for k, c in counts.items():
    if k.startswith("void cub::") or k.startswith("cub::"):
        k = "cub::" + k.split("::")[-1][:36]
    print("%-44s" % k[:43] + "".join("%10d" % c[x] for x in cols))
    tot.update(c)
print("%-44s" % "TOTAL" + "".join("%10d" % tot[x] for x in cols))
# the TMA + mbarrier sequence of k_seedw, verbatim
print("\n# the staging sequence of k_seedw<8,true,8,2> (first occurrence of each)")
seen = set(); on = False
for line in out.splitlines():
    if "Function :" in line:
        on = "k_seedwILi8ELb1ELi8ELi2E" in line
    if on:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
        if m and any(t in m.group(1) for t in ("UBLKCP", "SYNCS", "ELECT")):
            key = m.group(1).split()[0] if not m.group(1).startswith("@") else m.group(1).split()[1]
            if key not in seen:
                seen.add(key); print("   " + m.group(1).strip())
