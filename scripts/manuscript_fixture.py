#!/usr/bin/env python
"""The reference's manuscript data set through both binaries: 21 Enterococcus genomes
(61 Mbp) as the database, their annotated genes (48 .. 6,978 bp, budgets up to 139 at -i 0.98) as queries, ALLPATHS.  Long
queries with large budgets are the shape that takes the Myers prefix filter (k_filter) and the widest band classes / the
generic global-scratch band of k_extend -- nothing in the synthetic tiers does.

The two zips are the reference repository's manuscript/21Genomes.zip and manuscript/Genes21Genomes.zip; they are data, not source,
and are NOT committed: copy them to tests/fixtures_local/ (git-ignored, travels to the GPU box) before running.  The published
allpaths.b6 is not part of the reference checkout, so the golden output is the UNMODIFIED reference binary (oracle/_ref/burst15) run
on the same files here.  Prints one JSON line."""
import argparse, json, os, subprocess, sys, time, zipfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(cmd, cwd, timeout=None):
    t0 = time.time()
    try:
        r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return None, "timeout after %d s" % timeout
    run.last_stderr = r.stderr
    if r.returncode != 0:
        raise SystemExit("FAILED (%d): %s\n%s\n%s" % (r.returncode, " ".join(cmd), r.stdout[-3000:], r.stderr[-3000:]))
    return time.time() - t0, r.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--every", type=int, default=3, help="use every n-th gene (1 = all 61,693)")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--ident", default="0.98")
    ap.add_argument("--ref-timeout", type=int, default=1200)
    ap.add_argument("--dir", default="/tmp/ms")
    ap.add_argument("--skip-reference", action="store_true")
    a = ap.parse_args()
    fx = os.path.join(ROOT, "tests", "fixtures_local")
    if not os.path.exists(os.path.join(fx, "21Genomes.zip")):
        raise SystemExit("tests/fixtures_local/21Genomes.zip is missing (see the docstring)")
    os.makedirs(a.dir, exist_ok=True)
    for z in ("21Genomes.zip", "Genes21Genomes.zip"):
        zipfile.ZipFile(os.path.join(fx, z)).extractall(a.dir)
    lines = open(os.path.join(a.dir, "combined.fixed.fna")).read().splitlines()
    n = 0
    with open(os.path.join(a.dir, "genes.fna"), "w") as f:
        for i in range(0, len(lines) - 1, 2):
            if (i // 2) % a.every == 0:
                f.write(lines[i] + "\n" + lines[i + 1] + "\n"); n += 1
    ours = os.path.join(ROOT, "burst_b200", "host", "burst-b200"); ref = os.path.join(ROOT, "oracle", "_ref", "burst15")
    mk, _ = run([ours, "-r", "Concat.fasta", "-d", "DNA", "7000", "-o", "ms.edx", "-a", "ms.acx", "-s", "-i", a.ident, "--acx-n", "15"], a.dir)
    common = ["-r", "ms.edx", "-a", "ms.acx", "-q", "genes.fna", "-m", "ALLPATHS", "-i", a.ident, "--noprogress"]
    out = {"fixture": "manuscript 21 genomes x genes", "genes": n, "every": a.every, "threads": a.threads, "makedb_s": round(mk, 1), "flags": " ".join(common),
           "edx_bytes": os.path.getsize(os.path.join(a.dir, "ms.edx")), "acx_bytes": os.path.getsize(os.path.join(a.dir, "ms.acx"))}
    t, so = run([ours] + common + ["-o", "ours.b6", "-t", str(a.threads)], a.dir)
    out["ours_wall_s"] = round(t, 2)
    out["ours_stdout_tail"] = [l.strip() for l in so.splitlines() if "[Accel]" in l or "[time]" in l or "[engine]" in l][-14:]
    out["ours_stderr_tail"] = [l.strip() for l in run.last_stderr.splitlines() if "burst_b200" in l][-12:]
    rows = sorted(open(os.path.join(a.dir, "ours.b6"), "rb").read().splitlines())
    out["rows"] = len(rows)
    if a.skip_reference:
        print(json.dumps(out)); return
    t, sr = run([ref] + common + ["-o", "ref.b6", "-t", str(a.threads)], a.dir, timeout=a.ref_timeout)
    if t is None:
        out["reference"] = sr
    else:
        want = sorted(open(os.path.join(a.dir, "ref.b6"), "rb").read().splitlines())
        out.update(reference="burst15 -t %d" % a.threads, reference_wall_s=round(t, 2), reference_rows=len(want), b6_sorted_identical=rows == want, speedup_wall=round(t / out["ours_wall_s"], 2))
        if rows != want:
            sa, sb = set(rows), set(want)
            out["only_ours"] = len(sa - sb); out["only_reference"] = len(sb - sa)
            out["first_difference"] = [x.decode() for x in (sorted(sa - sb)[:2] + sorted(sb - sa)[:2])]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
