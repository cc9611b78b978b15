"""The synthetic bench workloads (burst_b200/synth.py) are self-consistent: strands sorted as the reference sorts them, runs =
(bunch, candidate) visits in order, the compact strand form describes the same batch, and the planted source of every read lies
within the budget (checked with the scalar oracle on a few reads)."""
import numpy as np
import pytest
from burst_b200 import synth


def check_form(w, qbunch=16):
    nq = len(w["qoff"]) - 1
    assert nq == 2 * w["n_reads"] and len(w["strand"]) == nq and len(w["slot"]) == nq
    # strands ascending under byte comparison
    prev = None
    for q in range(0, nq, max(1, nq // 400)):
        s = w["qcodes"][int(w["qoff"][q]):int(w["qoff"][q + 1])].tobytes()
        assert prev is None or prev <= s
        prev = s
    # runs follow the bunch -> candidate lists, in order, and the tasks are their expansion
    runs = w["runs"]; co = w["cand_off"].astype(np.int64)
    assert len(runs) == int(co[-1]) == len(w["cand"]) and np.array_equal(runs["clump"], w["cand"])
    b = np.repeat(np.arange(len(co) - 1), np.diff(co))
    assert np.array_equal(runs["query0"], b * qbunch) and np.array_equal(runs["nq"], np.minimum(qbunch, nq - b * qbunch))
    assert len(w["tasks"]) == int(runs["nq"].sum())
    # strand words: read index and orientation; the forward strand of read r equals the read as sequenced
    r = w["strand"] & 0x7FFFFFFF
    assert np.array_equal(r, w["slot"])
    q = int(np.nonzero((w["strand"] >> 31) == 0)[0][0]); rd = int(r[q])
    rl = w["rlen"].astype(np.int64); ro = np.concatenate([[0], np.cumsum(rl)])
    assert np.array_equal(w["qcodes"][int(w["qoff"][q]):int(w["qoff"][q + 1])], w["rcodes"][ro[rd]:ro[rd + 1]])
    # every strand that matches the database has the clump it was cut from among its bunch's candidates
    for q in np.nonzero(w["match"])[0][:: max(1, nq // 300)]:
        bb = q // qbunch
        assert w["true_clump"][w["slot"][q]] in w["cand"][co[bb]:co[bb + 1]]


def test_shotgun_workload_form():
    w = synth.bunch_workload(3000, 100, 2, 8 << 20, 214, seed=3)
    check_form(w)
    assert int(w["budget"].max()) == 2


def test_amplicon_workload_form_and_planted_reads(oracle):
    w = synth.amplicon_workload(1500, 292, 5, 2 << 20, 1400, seed=4, budget=9, halo=3)
    check_form(w)
    nq = len(w["qoff"]) - 1
    assert 3 <= len(w["tasks"]) / nq <= 40                                  # a few clump visits per strand at this size
    # the read's source lane is within the number of substitutions planted (<= 5 <= budget 9)
    S = oracle.score_table(1)
    qs = np.nonzero(w["match"])[0][:12]
    codes, qoff = synth.concat_queries([w["qcodes"][int(w["qoff"][q]):int(w["qoff"][q + 1])] for q in qs])
    tq = np.arange(len(qs), dtype=np.uint32); tc = w["true_clump"][w["slot"][qs]].astype(np.uint32)
    hits, best = oracle.run_tasks(w["packed"], w["clump_off"], w["clump_len"], codes, qoff, np.full(len(qs), 9, np.uint16),
                                  np.arange(len(qs), dtype=np.uint32), len(qs), tq, tc, S, 1)
    for i, q in enumerate(qs):
        h = hits[(hits["task"] == i) & (hits["lane"] == w["true_lane"][w["slot"][q]])]
        assert len(h) == 1 and int(h["ed"][0]) <= 5
