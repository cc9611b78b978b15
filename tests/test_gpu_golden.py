"""GPU: the CUDA path against the committed golden vectors of the reference's own kernels, and the
drop-in binary against the .b6 files the reference binary wrote (tests/golden/cli)."""
import json
import os
import subprocess
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
BIN = os.path.join(ROOT, "burst_b200", "host", "burst-b200")


@pytest.mark.parametrize("z", [0, 1])
def test_engine_matches_golden_reference_vectors(z):
    from burst_b200.engine import Engine, default_scoring, MODE_MIN
    g = np.load(os.path.join(GOLD, "kernels_z%d.npz" % z))
    eng = Engine(0)
    eng.set_scoring(default_scoring(z))
    checked = 0
    for i in range(len(g["clen"])):
        packed = g["packed"][g["packed_off"][i]:g["packed_off"][i + 1]]
        q = g["q"][g["q_off"][i]:g["q_off"][i + 1]]
        emac, rm = int(g["emac"][i]), int(g["min"][i])
        eng.load_db(packed, np.array([g["clen"][i]], np.uint32))
        hits, best = eng.align(q, np.array([0, len(q)], np.uint64), np.array([emac], np.uint16), None, MODE_MIN)
        if rm == 0xFFFFFFFF or rm > emac:
            assert len(hits) == 0 and best[0] == 0xFFFF, i
            continue
        assert best[0] == rm
        want = [(lane, int(g["mins"][i][lane]), int(g["gq"][i][lane]), int(g["gr"][i][lane]), int(g["fp"][i][lane]))
                for lane in range(16) if g["mins"][i][lane] == rm]
        got = [(int(h["lane"]), int(h["ed"]), int(h["gap_q"]), int(h["gap_r"]), int(h["final_pos"])) for h in hits]
        assert got == want, (i, got, want)
        checked += len(want)
    eng.close()
    assert checked > 60


CASES = sorted(os.listdir(os.path.join(GOLD, "cli")))


@pytest.mark.parametrize("case", CASES)
def test_cli_matches_reference_b6(case, tmp_path):
    cli_case(case, tmp_path, [])


@pytest.mark.parametrize("case", [c for c in CASES if c.startswith("acx_")])
def test_cli_device_candidates_matches_reference_b6(case, tmp_path):
    """--device-candidates: the accelerator on the GPU (bg_load_acx + bg_search_bunches_into, k_candgen); cases whose accelerated bin
    holds ambiguous queries fall back to the host lists by design and must say so"""
    out = cli_case(case, tmp_path, ["--device-candidates"])
    assert ("candidates on the device" in out) or ("--device-candidates not used" in out) or ("Using ACCELERATOR" not in out)
    if case in ("acx_allpaths_fr", "acx_best"):
        assert "candidates on the device" in out


def cli_case(case, tmp_path, extra):
    d = os.path.join(GOLD, "cli", case)
    meta = json.load(open(os.path.join(d, "case.json")))
    out = str(tmp_path / "out.b6")
    args = [out if a == "OUT" else a for a in meta["args"]]
    if "db.acx" in args:      # committed gzip-compressed (the k-mer length table is almost all zeros)
        import gzip, shutil
        acx = str(tmp_path / "db.acx")
        with gzip.open(os.path.join(d, "db.acx.gz"), "rb") as fi, open(acx, "wb") as fo:
            shutil.copyfileobj(fi, fo)
        args[args.index("db.acx")] = acx
    r = subprocess.run([BIN] + args + ["--noprogress"] + extra, cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = sorted(open(out).read().splitlines())
    want = sorted(open(os.path.join(d, "expected.b6")).read().splitlines())
    assert len(got) == len(want), (len(got), len(want))
    diff = [(a, b) for a, b in zip(got, want) if a != b]
    assert not diff, "%d rows differ, first: %s" % (len(diff), diff[0])
    return r.stdout
