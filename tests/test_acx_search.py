"""Candidate generation behind the C ABI (bg_load_acx + bg_search_bunches_into): the candidates of a bunch are
the reference's (burst.c:4085-4168: count above the bunch threshold, descending count, per-query skip, BadList), the alignment is
the usual one.  The CPU tier checks the oracle-backed stand-in against a Python restatement of the rule; the GPU tier checks the
CUDA path (k_candgen) against the stand-in on the same input, hit for hit."""
import os
import numpy as np
import pytest
from burst_b200 import synth
from burst_b200.engine import Engine, HIT_DTYPE, XHIT_DTYPE, RUN_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "oracle", "_sim", "libburst_b200_sim.so")


def case(seed, n_reads=150, family=False, N=12):
    rng = np.random.default_rng(seed)
    if family:
        refs = synth.mutation_tree_refs(rng, 16 * 30, 420, levels=(0.25, 0.10, 0.05, 0.01))
        refs.sort(key=lambda r: r.tobytes())
    else:
        refs = synth.random_refs(16 * 40, 260, rng, jitter=8)
        for i in range(0, 16 * 20, 37):                                   # a few references with a near copy in another clump
            refs[16 * 20 + i] = synth.mutate(refs[i], 1, rng)
    packed, off, clen = synth.pack_clumps(refs)
    reads, origin = synth.reads_from_clumps(packed, off, clen, n_reads, 120 if family else 100, 2, rng, rc_rate=0.5)
    reads = [r[:int(rng.integers(60, len(r) + 1))] for r in reads]
    reads[3] = reads[3][:10]                                              # shorter than a word: no candidates of its own
    bad = [5, len(clen) - 2]
    lens, post, badl, lists = synth.build_acx(refs, N, bad=bad)
    return rng, refs, packed, off, clen, reads, lens, post, badl, lists


def per_strand(h):
    """hits grouped by strand, inside a strand in arrival order (the order the host folds them in)"""
    o = np.argsort(h["query"], kind="stable")
    return h[o]


def run_search(e, packed, clen, reads, budgets, lens, post, badl, mode, N=12, heuristic=False, skip_bad=False):
    B = synth.strand_batch(reads, budgets, 16, lambda b, rd, rc: [])
    e.load_db(packed, clen)
    e.load_acx(lens, post, N, False, badl)
    buf = np.zeros(200000, XHIT_DTYPE); best = np.full(B["nreads"], 0xFFFF, np.uint16)
    n = e.search_bunches_into(Engine.pack2(B["rcodes"]), B["rlen"], B["rbudget"], B["strand"], 16, buf, best, mode, heuristic=heuristic, skip_bad=skip_bad)
    return B, per_strand(buf[:n].copy()), best


@pytest.mark.parametrize("family", [False, True])
def test_standin_candidates_follow_the_rule(oracle, family):
    if not os.path.exists(SIM):
        pytest.skip("oracle/_sim not built")
    rng, refs, packed, off, clen, reads, lens, post, badl, lists = case(5 + family, n_reads=60, family=family)
    budgets = [oracle.budget(0.97 if family else 0.98, len(r)) for r in reads]
    e = Engine(0, lib_path=SIM)
    try:
        for heuristic, skip_bad in ((False, False), (True, True)):
            B, got, best = run_search(e, packed, clen, reads, budgets, lens, post, badl, 0, heuristic=heuristic, skip_bad=skip_bad)
            strands = [B["qcodes"][int(B["qoff"][i]):int(B["qoff"][i + 1])] for i in range(len(B["strand"]))]
            runs = synth.acx_runs(strands, B["budget"], 16, 12, lists, badl, len(clen), heuristic=heuristic, skip_bad=skip_bad)
            assert len(runs) > len(strands) // 16
            R = np.array(runs, np.uint32).reshape(-1, 3).view(RUN_DTYPE).reshape(-1)
            hits, best2 = e.align(B["qcodes"], B["qoff"], B["budget"], None, 0, slot=B["slot"], nslots=B["nreads"], runs=R)
            want = np.zeros(len(hits), XHIT_DTYPE)
            want["query"] = R["query0"][hits["task"] >> 4] + (hits["task"] & 15); want["clump"] = R["clump"][hits["task"] >> 4]
            for f in ("lane", "ed", "gap_q", "gap_r", "final_pos"):
                want[f] = hits[f]
            assert np.array_equal(per_strand(want), got), (heuristic, skip_bad, len(want), len(got))
            assert np.array_equal(best, best2) and len(got) > 20
    finally:
        e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("family", [False, True])
@pytest.mark.parametrize("mode", [0, 1])
def test_device_candidates_match_the_standin(oracle, family, mode):
    rng, refs, packed, off, clen, reads, lens, post, badl, lists = case(40 + family + 2 * mode, n_reads=400, family=family)
    budgets = [oracle.budget(0.97 if family else 0.98, len(r)) for r in reads]
    sim = Engine(0, lib_path=SIM); dev = Engine(0)
    try:
        for heuristic, skip_bad in ((False, False), (True, True)):
            _, want, wbest = run_search(sim, packed, clen, reads, budgets, lens, post, badl, mode, heuristic=heuristic, skip_bad=skip_bad)
            _, got, gbest = run_search(dev, packed, clen, reads, budgets, lens, post, badl, mode, heuristic=heuristic, skip_bad=skip_bad)
            assert len(want) > 100
            assert np.array_equal(gbest, wbest)
            assert len(got) == len(want) and np.array_equal(got, want), (family, mode, heuristic, skip_bad)
    finally:
        sim.close(); dev.close()
