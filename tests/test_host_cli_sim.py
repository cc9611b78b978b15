"""CPU-only: the HOST logic of the drop-in binary (CLI, FASTA/.edx/.acx readers, query preprocessing,
candidate generation, pod lists, BEST/ALLPATHS/CAPITALIST/FORAGE reporters, taxonomy) against the
.b6 files the reference binary wrote (tests/golden/cli, made by scripts/make_golden.py).

The binary under test is oracle/_sim/burst-b200-sim: burst_b200/host/burst_b200.c linked against
oracle/abi_sim.c, a stand-in for the engine ABI backed by the scalar oracle -- test infrastructure
that exists so this tier can run without a GPU.  The same cases run against the real binary and
the CUDA engine in tests/test_gpu_golden.py."""
import gzip
import json
import os
import shutil
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "cli")
SIM = os.path.join(ROOT, "oracle", "_sim", "burst-b200-sim")
CASES = sorted(os.listdir(GOLD))


@pytest.fixture(scope="module")
def sim():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "sim"], check=True)
    return SIM


def run_case(binary, case, tmp_path, extra=()):
    d = os.path.join(GOLD, case)
    meta = json.load(open(os.path.join(d, "case.json")))
    out = str(tmp_path / "out.b6")
    args = [out if a == "OUT" else a for a in meta["args"]]
    if "db.acx" in args:
        acx = str(tmp_path / "db.acx")
        with gzip.open(os.path.join(d, "db.acx.gz"), "rb") as fi, open(acx, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        args[args.index("db.acx")] = acx
    r = subprocess.run([binary] + args + ["--noprogress"] + list(extra), cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = sorted(open(out).read().splitlines())
    want = sorted(open(os.path.join(d, "expected.b6")).read().splitlines())
    return got, want


@pytest.mark.parametrize("case", CASES)
def test_host_logic_matches_reference_b6(sim, case, tmp_path):
    got, want = run_case(sim, case, tmp_path)
    assert len(got) == len(want), (len(got), len(want))
    diff = [(a, b) for a, b in zip(got, want) if a != b]
    assert not diff, "%d rows differ, first: %s" % (len(diff), diff[0])


@pytest.mark.parametrize("case", ["acx_allpaths_fr", "acx_best", "acx_capitalist_tax_iupac"])
def test_device_candidate_driver_matches_reference_b6(sim, case, tmp_path):
    """--device-candidates: the host side of the path (distinct reads of a batch at 2 bits per base, strand words, hits that name
    strand and clump) over the stand-in's CPU statement of the candidate rule; the third case must fall back to the host lists"""
    got, want = run_case(sim, case, tmp_path, extra=["--device-candidates"])
    assert got == want


@pytest.mark.parametrize("threads", ["7", "64"])
def test_bunch_size_does_not_change_rows(sim, tmp_path, threads):
    """-t only changes QBUNCH (burst.c:4019-4021), never the reported rows."""
    got, want = run_case(sim, "acx_allpaths_fr", tmp_path, extra=["-t", threads])
    assert got == want


@pytest.mark.parametrize("case", CASES)
def test_rows_formatted_by_the_thread_team_in_file_order(sim, case, tmp_path, monkeypatch):
    """Every reporter (BEST, ALLPATHS, FORAGE, CAPITALIST after its global tally) formats blocks of queries on all threads and writes the
    blocks in query order: with blocks of 7 queries and 5 threads the FILE must be byte-identical (not just the same set of rows) to the
    file the same 5-thread run writes through one block (the sequential loop), and hold the reference's rows."""
    got_seq, want = run_case(sim, case, tmp_path, extra=["-t", "5"])
    raw_seq = open(str(tmp_path / "out.b6"), "rb").read()
    monkeypatch.setenv("BURST_B200_REPORT_BLOCK", "7")
    got_blk, _ = run_case(sim, case, tmp_path, extra=["-t", "5"])
    raw_blk = open(str(tmp_path / "out.b6"), "rb").read()
    assert raw_blk == raw_seq
    if case != "acx_forage_mixed_lengths":       # (FORAGE reports every lane of every clump a BUNCH visits: its rows depend on the bunch size, i.e. on -t, in the reference too)
        assert got_blk == want


def test_usage_errors_exit_codes(sim, tmp_path):
    r = subprocess.run([sim, "-r", "x.fa", "-q"], capture_output=True, text=True)
    assert r.returncode == 1
    # a missing reference is "invalid input file" -> 1 (burst.c:4894-4897); a missing query file -> 2 (burst.c:639)
    r = subprocess.run([sim, "-r", "/nonexistent.fa", "-q", "/nonexistent.fa", "-o", str(tmp_path / "o.b6")], capture_output=True, text=True)
    assert r.returncode == 1
    refs = os.path.join(GOLD, "fasta_best", "refs.fa")
    r = subprocess.run([sim, "-r", refs, "-q", "/nonexistent.fa", "-o", str(tmp_path / "o.b6")], capture_output=True, text=True)
    assert r.returncode == 2


def test_query_order_with_shared_prefixes_and_the_thread_team(sim, tmp_path):
    """Query preprocessing (burst.c:2980-3223) sorts code strings by strcmp; the host sorts (16-base key, index) pairs on all threads and
    only follows the pointers when keys tie.  70 k reads whose first 17 bases come from 50 prefixes (every comparison inside a family ties on
    the key), reads that are prefixes of other reads, and duplicates: the rows of a BEST run must come out in strcmp order of the code
    strings with duplicates in input order, and -t 1 (plain qsort) and -t 8 (sorted pieces + merges) must write the same file."""
    import numpy as np
    rng = np.random.default_rng(99)
    code = {"A": 1, "C": 2, "G": 3, "T": 4}
    pre = ["".join("ACGT"[i] for i in rng.integers(0, 4, 17)) for _ in range(50)]
    uniq = set()
    while len(uniq) < 34000:
        uniq.add(pre[int(rng.integers(0, 50))] + "".join("ACGT"[i] for i in rng.integers(0, 4, int(rng.integers(3, 8)))))
    uniq = sorted(uniq)
    reads = list(uniq) + [u[:-2] for u in uniq[:2000]] + [uniq[int(i)] for i in rng.integers(0, len(uniq), 36000)]
    order = rng.permutation(len(reads))
    reads = [reads[i] for i in order]
    assert len(reads) >= 70000
    with open(tmp_path / "q.fa", "w") as f:
        f.write("".join(">r%d\n%s\n" % (i, r) for i, r in enumerate(reads)))
    with open(tmp_path / "r.fa", "w") as f:
        f.write("".join(">ref%d\n%s\n" % (i, "".join("ACGT"[j] for j in rng.integers(0, 4, 30))) for i in range(16)))
    files = []
    for t in ("1", "8"):
        out = str(tmp_path / ("o%s.b6" % t))
        r = subprocess.run([sim, "-r", str(tmp_path / "r.fa"), "-q", str(tmp_path / "q.fa"), "-o", out, "-m", "BEST", "-i", "0.5", "-t", t, "--noprogress"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        files.append(open(out, "rb").read())
    assert files[0] == files[1]
    names = [l.split(b"\t", 1)[0].decode() for l in files[0].splitlines()]
    keyed = sorted(range(len(reads)), key=lambda i: (bytes(code[c] for c in reads[i]), i))
    assert names == ["r%d" % i for i in keyed]
    # both strands (-fr): 2 x 36 k unique strands, the merged sort of forward and reverse-complement strings (burst.c:3178-3186)
    both = []
    for t in ("1", "8"):
        out = str(tmp_path / ("f%s.b6" % t))
        r = subprocess.run([sim, "-r", str(tmp_path / "r.fa"), "-q", str(tmp_path / "q.fa"), "-o", out, "-m", "BEST", "-i", "0.5", "-fr", "-t", t, "--noprogress"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        both.append(open(out, "rb").read())
    assert both[0] == both[1] and len(both[0].splitlines()) == len(reads)
