"""CPU-only: the .edx / .acx WRITER of the drop-in binary (-d; layouts burst.c:2758-2839, 3501-3530).

For every golden case that uses a database: build the database from the case's refs.fa with OUR -d, then
  * the UNMODIFIED reference binary (oracle/_ref/burst12, when present) must load both files and align the case's
    queries on them, and
  * our host driver (oracle-backed stand-in engine: this tier has no GPU) must produce the same sorted rows on the same
    files -- and, in BEST mode, the rows of the golden file, which the reference wrote from a database it built itself
    (a different shearing; BEST rows do not depend on clump composition).
The same round trip runs against the CUDA engine in tests/test_gpu_golden.py."""
import json
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "cli")
SIM = os.path.join(ROOT, "oracle", "_sim", "burst-b200-sim")
REF12 = os.path.join(ROOT, "oracle", "_ref", "burst12")
DB_CASES = [c for c in sorted(os.listdir(GOLD)) if json.load(open(os.path.join(GOLD, c, "case.json"))).get("make_edx")]


def build_and_run(builder, aligner, case, tmp_path, tag):
    d = os.path.join(GOLD, case)
    meta = json.load(open(os.path.join(d, "case.json")))
    edx, acx, out = str(tmp_path / "my.edx"), str(tmp_path / "my.acx"), str(tmp_path / (tag + ".b6"))
    if not os.path.exists(edx):
        mk = [builder, "-r", "refs.fa", "-o", edx] + meta["make_edx"] + (["-a", acx] if "db.acx" in meta["args"] else [])
        r = subprocess.run(mk, cwd=d, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    args = [{"OUT": out, "db.edx": edx, "db.acx": acx}.get(a, a) for a in meta["args"]]
    r = subprocess.run([aligner] + args + ["--noprogress", "-t", "1"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    return sorted(open(out).read().splitlines()), meta


@pytest.fixture(scope="module")
def sim():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "sim"], check=True)
    return SIM


@pytest.mark.parametrize("case", DB_CASES)
def test_written_db_round_trips_through_the_reference(sim, case, tmp_path):
    ours, meta = build_and_run(sim, sim, case, tmp_path, "ours")
    assert len(ours) > 50
    if os.path.exists(REF12):
        theirs, _ = build_and_run(sim, REF12, case, tmp_path, "ref")
        assert len(ours) == len(theirs)
        diff = [(a, b) for a, b in zip(ours, theirs) if a != b]
        assert not diff, "%d rows differ on the same written DB, first: %s" % (len(diff), diff[0])
    if "BEST" in meta["args"]:
        want = sorted(open(os.path.join(GOLD, case, "expected.b6")).read().splitlines())
        assert ours == want


def test_edx_header_fields(sim, tmp_path):
    import struct
    d = os.path.join(GOLD, "acx_best")
    edx, acx = str(tmp_path / "a.edx"), str(tmp_path / "a.acx")
    subprocess.run([sim, "-r", "refs.fa", "-o", edx, "-a", acx, "-d", "DNA", "140", "-s", "1", "-i", "0.97"], cwd=d, check=True, capture_output=True)
    b = open(edx, "rb").read()
    assert b[0] == (1 << 7 | 1 << 6 | 3)                                   # burst.c:2836: edx, sheared, no fingerprints, version 3
    heads, shear, totR, origTotR, nclumps, maxLenR = struct.unpack_from("<QIIIII", b, 1)
    assert shear == int(140 / 0.97) and totR == origTotR and nclumps == (totR + 15) // 16 and maxLenR <= 2 * shear
    a = open(acx, "rb").read(5)
    assert a[0] == (1 << 7 | 1 << 6 | 0)                                   # burst.c:3501: acx, built with N penalised, small format
    assert os.path.getsize(acx) >= 5 + 4 * (1 << 24)
