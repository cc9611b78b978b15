#!/usr/bin/env python
"""Generate tests/golden/ from the UNMODIFIED reference (run here, where /root/reference exists).

  kernels_z{0,1}.npz   random (query, clump) tasks with the outputs of the reference's own
                       aded_mat16L / aded_mat16 / reScoreM_mat16 (oracle/_ref/libburstref.so)
  cli/<case>/          small FASTA / taxonomy / .edx inputs, the command-line flags, and the .b6
                       the reference binary (oracle/_ref/burst12, -t 1) wrote for them

The fixtures are committed; this script is the record of how they were made.
"""
import json
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from burst_b200 import synth  # noqa: E402
from oracle.pyoracle import Reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
BURST12 = os.path.join(ROOT, "oracle", "_ref", "burst12")


def kernel_vectors(z, seed, n=160):
    rng = np.random.default_rng(seed)
    ref = Reference(); ref.set_scoring(z)
    recs = dict(packed=[], clen=[], q=[], emac=[], variant=[], min=[], mins=[], score=[], fp=[], gr=[], gq=[])
    for it in range(n):
        qlen = int(rng.integers(10, 150)); clen = int(rng.integers(qlen + 4, 300))
        refs = synth.random_refs(16, clen, rng, jitter=(clen // 5 if it % 2 else 0), iupac_rate=(0.02 if it % 3 == 0 else 0))
        packed, off, clens = synth.pack_clumps(refs)
        reads, _ = synth.reads_from_clumps(packed, off, clens, 1, qlen, int(rng.integers(0, 6)), rng)
        q = reads[0]
        if it % 4 == 0:
            m = rng.random(len(q)) < 0.03; q[m] = rng.integers(5, 16, int(m.sum()), dtype=np.uint8)
        emac = int(rng.integers(0, 8)); variant = it & 1
        mn, mins, score, fp, gr, gq = ref.task(packed, int(clens[0]), q, emac, variant)
        recs["packed"].append(packed); recs["clen"].append(int(clens[0])); recs["q"].append(q); recs["emac"].append(emac)
        recs["variant"].append(variant); recs["min"].append(mn); recs["mins"].append(mins); recs["score"].append(score)
        recs["fp"].append(fp); recs["gr"].append(gr); recs["gq"].append(gq)
    ref.set_scoring(1)
    np.savez_compressed(os.path.join(GOLD, "kernels_z%d.npz" % z),
                        packed=np.concatenate(recs["packed"]), packed_off=np.cumsum([0] + [len(p) for p in recs["packed"]]),
                        clen=np.array(recs["clen"]), q=np.concatenate(recs["q"]), q_off=np.cumsum([0] + [len(p) for p in recs["q"]]),
                        emac=np.array(recs["emac"]), variant=np.array(recs["variant"]), min=np.array(recs["min"], np.uint32),
                        mins=np.stack(recs["mins"]), score=np.stack(recs["score"]), fp=np.stack(recs["fp"]), gr=np.stack(recs["gr"]),
                        gq=np.stack(recs["gq"]), z=z)


def write_tax(path, names, rng):
    with open(path, "w") as f:
        for n in names:
            k = ["k__K%d" % rng.integers(0, 2), "p__P%d" % rng.integers(0, 3), "c__C%d" % rng.integers(0, 3),
                 "o__O%d" % rng.integers(0, 4), "f__F%d" % rng.integers(0, 4), "g__G%d" % rng.integers(0, 6), "s__S%d" % rng.integers(0, 9)]
            f.write("%s\t%s\n" % (n.split()[0] if False else n, ";".join(k)))


def family_refs(rng, nfam, per, length, div):
    """references in families of near-identical members, so that ties / ALLPATHS / CAPITALIST have work"""
    out = []
    for f in range(nfam):
        base = rng.integers(1, 5, length + int(rng.integers(-30, 31)), dtype=np.uint8)
        for m in range(per):
            out.append(synth.mutate(base, int(div * len(base) * rng.random()), rng) if m else base.copy())
    return out


def make_reads(rng, refs, n, lo, hi, max_err_frac, iupac=0.0, dup=0.1):
    reads = []
    for i in range(n):
        if reads and rng.random() < dup:
            reads.append(reads[int(rng.integers(0, len(reads)))].copy()); continue
        r = refs[int(rng.integers(0, len(refs)))]
        L = int(rng.integers(lo, hi + 1)); L = min(L, len(r))
        o = int(rng.integers(0, len(r) - L + 1))
        q = synth.mutate(r[o:o + L], int(rng.integers(0, int(max_err_frac * L) + 1)), rng)
        if rng.random() < 0.5 and i % 3 == 0:
            q = synth.RC_TABLE[q[::-1]]
        if iupac:
            m = rng.random(len(q)) < iupac; q[m] = rng.integers(5, 16, int(m.sum()), dtype=np.uint8)
        reads.append(q.astype(np.uint8))
    # a few unalignable reads
    for i in range(max(2, n // 50)):
        reads.append(rng.integers(1, 5, lo, dtype=np.uint8))
    return reads


def run_ref(args, cwd):
    r = subprocess.run([BURST12] + args + ["-t", "1", "--noprogress"], cwd=cwd, capture_output=True, text=True)
    return r.returncode, r.stdout + r.stderr


def cli_case(name, seed, flags, nfam=10, per=5, length=420, nreads=260, lo=90, hi=130, err=0.04, iupac=0.0, tax=False,
             edx=None, multiline=False, acx=False):
    rng = np.random.default_rng(seed)
    d = os.path.join(GOLD, "cli", name)
    shutil.rmtree(d, ignore_errors=True); os.makedirs(d)
    refs = family_refs(rng, nfam, per, length, 0.03)
    if iupac:
        for r in refs[::4]:
            m = rng.random(len(r)) < 0.004; r[m] = rng.integers(5, 16, int(m.sum()), dtype=np.uint8)
    rnames = ["ref%03d fam%d member %d" % (i, i // per, i % per) for i in range(len(refs))]
    with open(os.path.join(d, "refs.fa"), "w") as f:
        for n, s in zip(rnames, refs):
            txt = "".join(synth.ALPHABET[int(c)] for c in s)
            if multiline:
                txt = "\n".join(txt[i:i + 70] for i in range(0, len(txt), 70))
            f.write(">%s\n%s\n" % (n, txt))
    reads = make_reads(rng, refs, nreads, lo, hi, err, iupac=iupac)
    qnames = ["q%04d extra words" % i for i in range(len(reads))]
    synth.to_fasta(os.path.join(d, "queries.fa"), qnames, reads)
    args = []
    if tax:
        write_tax(os.path.join(d, "tax.txt"), rnames, rng)
    ref_arg = "refs.fa"
    acx_args = []
    if edx:
        rc, log = run_ref(["-r", "refs.fa", "-o", "db.edx"] + edx + (["-a", "db.acx"] if acx else []), d)
        assert rc == 0 and os.path.exists(os.path.join(d, "db.edx")), log
        ref_arg = "db.edx"
        if acx:   # the 4^12-entry length table is almost all zeros: keep the accelerator gzip-compressed
            import gzip
            with open(os.path.join(d, "db.acx"), "rb") as fi, gzip.open(os.path.join(d, "db.acx.gz"), "wb", 9) as fo:
                shutil.copyfileobj(fi, fo)
            acx_args = ["-a", "db.acx"]
    args = ["-r", ref_arg, "-q", "queries.fa", "-o", "expected.b6"] + acx_args + flags + (["-b", "tax.txt"] if tax else [])
    rc, log = run_ref(args, d)
    assert rc == 0, log
    rows = open(os.path.join(d, "expected.b6")).read().splitlines()
    json.dump({"args": [a if a != "expected.b6" else "OUT" for a in args], "rows": len(rows), "make_edx": edx,
               "reference": "oracle/_ref/burst12 (gcc -O3 build of /root/reference/burst.c, -t 1)"},
              open(os.path.join(d, "case.json"), "w"), indent=1)
    print("%-28s %5d rows  %s" % (name, len(rows), " ".join(args)))
    assert len(rows) > 20, log
    if acx:
        os.remove(os.path.join(d, "db.acx"))
        return log


def main():
    os.makedirs(GOLD, exist_ok=True)
    kernel_vectors(1, 101); kernel_vectors(0, 202)
    cli_case("fasta_best", 1, ["-m", "BEST", "-i", "0.97"])
    cli_case("fasta_best_multiline_refs", 2, ["-m", "BEST", "-i", "0.95"], multiline=True, lo=60, hi=200)
    cli_case("fasta_allpaths_fr", 3, ["-m", "ALLPATHS", "-i", "0.95", "-fr"])
    cli_case("fasta_capitalist_tax", 4, ["-m", "CAPITALIST", "-i", "0.97"], tax=True)
    cli_case("fasta_capitalist_default", 5, ["-fr"])
    cli_case("fasta_forage_y_iupac", 6, ["-m", "FORAGE", "-i", "0.96", "-y"], iupac=0.004)
    cli_case("fasta_best_iupac_fr_tax", 7, ["-m", "BEST", "-i", "0.93", "-fr", "-bs"], iupac=0.004, tax=True)
    cli_case("fasta_allpaths_whitespace", 8, ["-m", "ALLPATHS", "-i", "0.98", "-w"])
    cli_case("edx_best", 9, ["-m", "BEST", "-i", "0.97"], length=900, edx=["-d", "DNA", "140", "-s", "1", "-i", "0.97"])
    cli_case("edx_allpaths_fr", 10, ["-m", "ALLPATHS", "-i", "0.97", "-fr"], length=900, edx=["-d", "DNA", "140", "-s", "1", "-i", "0.97"])
    cli_case("edx_capitalist_tax", 11, ["-m", "CAPITALIST", "-i", "0.97"], length=900, tax=True, edx=["-d", "DNA", "140", "-s", "1", "-i", "0.97"])
    cli_case("edx_quick_forage", 12, ["-m", "FORAGE", "-i", "0.96"], edx=["-d", "QUICK"])
    DB = ["-d", "DNA", "140", "-s", "1", "-i", "0.97"]
    cli_case("acx_best", 13, ["-m", "BEST", "-i", "0.97"], length=900, edx=DB, acx=True)
    cli_case("acx_allpaths_fr", 14, ["-m", "ALLPATHS", "-i", "0.97", "-fr"], length=900, edx=DB, acx=True)
    cli_case("acx_capitalist_tax_iupac", 15, ["-m", "CAPITALIST", "-i", "0.97", "-fr"], length=900, tax=True, iupac=0.003, edx=DB, acx=True)
    cli_case("acx_forage_mixed_lengths", 16, ["-m", "FORAGE", "-i", "0.95", "-fr"], length=900, lo=30, hi=130, iupac=0.003, edx=DB, acx=True)
    cli_case("acx_best_lowid_fallback_bin", 17, ["-m", "BEST", "-i", "0.90"], length=900, lo=60, hi=130, err=0.08, edx=["-d", "DNA", "140", "-s", "1", "-i", "0.90"], acx=True)
    cli_case("acx_y_wildcard", 18, ["-m", "ALLPATHS", "-i", "0.96", "-y"], length=900, iupac=0.004, edx=["-d", "DNA", "140", "-s", "1", "-i", "0.96", "-y"], acx=True)


if __name__ == "__main__":
    main()
