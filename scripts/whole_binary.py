#!/usr/bin/env python
"""Whole-program comparison on files both binaries read: the UNMODIFIED reference (oracle/_ref/burst12|burst15, built from
/root/reference/burst.c by oracle/Makefile) against the drop-in binary burst_b200/host/burst-b200, same .edx/.acx, same
reads, same flags.  Sorted .b6 must be byte-identical; wall times of both are reported (the reference with all host threads).

  --shape shotgun   uniform-random references (--mbp), 100 bp reads with exactly 2 edits (the LLsim model), DB15, -m BEST -i 0.98 -fr
                    (BASELINE.md section 2 "shotgun shape", BASELINE.json configs[1] at reduced size)
  --shape amplicon  mutation-tree references (25/10/5/3 % divergence), 292 bp reads from a jittered fixed start with 0-5 edits,
                    DB12, -m CAPITALIST -i 0.97 with a taxonomy table (BASELINE.json configs[2] at reduced size)

The database is written by burst-b200 -d (tests/test_makedb.py checks that the reference loads such files and finds the same rows).
Runs on the GPU box (needs a CUDA device for burst-b200); prints one JSON line."""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from burst_b200 import synth  # noqa: E402

ALPHA = np.frombuffer(b".ACGTNKMRYSWBVHD", np.uint8)


def write_fasta(path, names, seqs, width=0):
    with open(path, "wb") as f:
        for n, s in zip(names, seqs):
            f.write(b">" + n.encode() + b"\n" + ALPHA[s].tobytes() + b"\n")


def shotgun(args, d):
    rng = np.random.default_rng(args.seed)
    nref = max(1, args.mbp)                                       # 1 Mbp references
    refs = [rng.integers(1, 5, 1_000_000, dtype=np.uint8) for _ in range(nref)]
    write_fasta(os.path.join(d, "refs.fa"), ["ref%05d" % i for i in range(nref)], refs)
    n = args.reads
    cat = np.concatenate(refs)
    r = rng.integers(0, nref, n); st = rng.integers(0, 1_000_000 - 110, n)
    w = cat[(r * 1_000_000 + st)[:, None] + np.arange(100)[None, :]]          # n x 100 windows
    ins = np.zeros((n, 100), np.uint8)
    pos = np.argsort(rng.random((n, 100), dtype=np.float32), axis=1)[:, :2]    # exactly 2 edits at distinct positions (embalmlets/LLsim.c)
    typ = rng.integers(0, 5, (n, 2))
    rows = np.repeat(np.arange(n), 2); cols = pos.reshape(-1); t = typ.reshape(-1)
    sub = t < 3
    w[rows[sub], cols[sub]] = ((w[rows[sub], cols[sub]].astype(np.int64) - 1 + 1 + t[sub]) % 4 + 1).astype(np.uint8)
    w[rows[t == 3], cols[t == 3]] = 0
    ins[rows[t == 4], cols[t == 4]] = rng.integers(1, 5, int((t == 4).sum()), dtype=np.uint8)
    both = np.stack([ins, w], axis=2).reshape(n, -1)
    flip = rng.random(n) < 0.5
    reads = []
    for i in range(n):
        q = both[i][both[i] != 0]
        reads.append(synth.RC_TABLE[q[::-1]] if flip[i] else q)
    write_fasta(os.path.join(d, "reads.fa"), ["r%07d" % i for i in range(n)], reads)
    return dict(qlen=105, ident="0.98", mode="BEST", extra=["-fr"], acx_n=15, ref_bin="burst15")


def amplicon(args, d):
    rng = np.random.default_rng(args.seed)
    nrefs = max(64, int(args.mbp * 1e6 / 1400))
    # substitutions only (vectorised): root -> 4 phyla (25 %) -> 16 families (10 %) -> 64 genera (5 %) -> species (3 %)
    def diverge(s, frac):
        t = s.copy(); k = int(frac * len(s)); p = rng.choice(len(s), k, replace=False)
        t[p] = (t[p] - 1 + rng.integers(1, 4, k)) % 4 + 1
        return t.astype(np.uint8)
    nodes = [rng.integers(1, 5, 1400, dtype=np.uint8)]
    for lv in (0.25, 0.10, 0.05):
        nodes = [diverge(p, lv) for p in nodes for _ in range(4)]
    per = (nrefs + len(nodes) - 1) // len(nodes)
    refs, tax = [], []
    for gi, g in enumerate(nodes):
        for k in range(per):
            if len(refs) < nrefs:
                refs.append(diverge(g, 0.03 * rng.random()))
                tax.append("k__B;p__P%d;c__C%d;o__O%d;f__F%d;g__G%d;s__S%d" % (gi // 16, gi // 16, gi // 4, gi // 4, gi, len(refs)))
    names = ["otu%06d" % i for i in range(len(refs))]
    write_fasta(os.path.join(d, "refs.fa"), names, refs)
    with open(os.path.join(d, "tax.txt"), "w") as f:
        for n, t in zip(names, tax):
            f.write("%s\t%s\n" % (n, t))
    n = args.reads
    src = rng.integers(0, len(refs), n)
    reads = []
    for i in range(n):
        o = 60 + int(rng.integers(-5, 6))
        q = refs[src[i]][o:o + 292].copy()
        e = int(rng.integers(0, 6)); p = rng.choice(292, e, replace=False)
        q[p] = (q[p] - 1 + rng.integers(1, 4, e)) % 4 + 1
        reads.append(q.astype(np.uint8))
    write_fasta(os.path.join(d, "reads.fa"), ["a%07d" % i for i in range(n)], reads)
    return dict(qlen=320, ident="0.97", mode="CAPITALIST", extra=["-b", "tax.txt"], acx_n=12, ref_bin="burst12")


def run(cmd, cwd):
    t0 = time.time()
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    dt = time.time() - t0
    if r.returncode != 0:
        raise SystemExit("FAILED (%d): %s\n%s\n%s" % (r.returncode, " ".join(cmd), r.stdout[-3000:], r.stderr[-3000:]))
    return dt, r.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="shotgun", choices=["shotgun", "amplicon"])
    ap.add_argument("--mbp", type=int, default=100)
    ap.add_argument("--reads", type=int, default=200000)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--dir", default="/tmp/wb")
    ap.add_argument("--skip-reference", action="store_true")
    ap.add_argument("--ours", default=None, help="binary under test (default burst_b200/host/burst-b200)")
    ap.add_argument("--reuse", action="store_true", help="keep the files of an earlier run in --dir (same shape and sizes)")
    ap.add_argument("--ours-extra", default="", help="extra flags for the binary under test, e.g. --device-candidates")
    ap.add_argument("--acx-n", type=int, default=0, help="override the accelerator word length (12 or 15)")
    args = ap.parse_args()
    d = os.path.join(args.dir, args.shape); os.makedirs(d, exist_ok=True)
    t0 = time.time()
    have = args.reuse and os.path.exists(os.path.join(d, "db.edx")) and os.path.exists(os.path.join(d, "reads.fa"))
    if have:
        P = (dict(qlen=105, ident="0.98", mode="BEST", extra=["-fr"], acx_n=15, ref_bin="burst15") if args.shape == "shotgun"
             else dict(qlen=320, ident="0.97", mode="CAPITALIST", extra=["-b", "tax.txt"], acx_n=12, ref_bin="burst12"))
    else:
        P = shotgun(args, d) if args.shape == "shotgun" else amplicon(args, d)
    gen_s = time.time() - t0
    ours = args.ours or os.path.join(ROOT, "burst_b200", "host", "burst-b200")
    if args.acx_n:
        P["acx_n"] = args.acx_n; P["ref_bin"] = "burst%d" % args.acx_n
    ref = os.path.join(ROOT, "oracle", "_ref", P["ref_bin"])
    mk = 0.0
    if not have:
        mk, _ = run([ours, "-r", "refs.fa", "-d", "DNA", str(P["qlen"]), "-o", "db.edx", "-a", "db.acx", "-s", "1", "-i", P["ident"], "--acx-n", str(P["acx_n"]), "-t", str(args.threads)], d)
    common = ["-r", "db.edx", "-a", "db.acx", "-q", "reads.fa", "-m", P["mode"], "-i", P["ident"], "--noprogress"] + P["extra"]
    out = {"shape": args.shape, "mbp": args.mbp, "reads": args.reads, "threads": args.threads, "gpus": args.gpus, "generate_s": round(gen_s, 1), "makedb_s": round(mk, 1),
           "edx_bytes": os.path.getsize(os.path.join(d, "db.edx")), "acx_bytes": os.path.getsize(os.path.join(d, "db.acx")), "flags": " ".join(common)}
    t_ours, so = run([ours] + common + ["-o", "ours.b6", "-t", str(args.threads), "--gpus", str(args.gpus)] + args.ours_extra.split(), d)
    out["ours_extra"] = args.ours_extra
    out["ours_wall_s"] = round(t_ours, 2); out["ours_reads_per_s"] = round(args.reads / t_ours)
    out["ours_stdout_tail"] = [l.strip() for l in so.splitlines() if "[Accel]" in l or "Alignment time" in l or "[time]" in l][-12:]
    rows = sorted(open(os.path.join(d, "ours.b6"), "rb").read().splitlines())
    out["rows"] = len(rows)
    if not args.skip_reference and os.path.exists(ref):
        t_ref, sr = run([ref] + common + ["-o", "ref.b6", "-t", str(args.threads)], d)
        want = sorted(open(os.path.join(d, "ref.b6"), "rb").read().splitlines())
        out["reference"] = P["ref_bin"] + " -t %d" % args.threads
        out["reference_wall_s"] = round(t_ref, 2); out["reference_reads_per_s"] = round(args.reads / t_ref)
        out["reference_rows"] = len(want)
        out["b6_sorted_identical"] = rows == want
        if rows != want:
            out["first_difference"] = next(((a.decode(), b.decode()) for a, b in zip(rows, want) if a != b), ("(length)", "%d vs %d" % (len(rows), len(want))))
        out["speedup_wall"] = round(t_ref / t_ours, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
