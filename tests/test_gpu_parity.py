"""GPU parity: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.
Bit-exact on every integer the reference reports: edit distance, numGapQ, numGapR, finalPos,
the set of reported (task, lane) pairs and the per-slot minima."""
import numpy as np
import pytest
from burst_b200 import synth
from burst_b200.engine import default_scoring

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from burst_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def all_vs_all(nq, nc):
    tq = np.tile(np.arange(nq, dtype=np.uint32), nc)
    tc = np.repeat(np.arange(nc, dtype=np.uint32), nq)
    return tq, tc


def check(eng, oracle, packed, off, clen, reads, budget, tasks=None, mode=0, slot=None, nslots=0, z=1, best=None):
    codes, qoff = synth.concat_queries(reads)
    nq = len(reads)
    budget = np.asarray(budget, np.uint16)
    if slot is None:
        slot = np.arange(nq, dtype=np.uint32); nslots = nq
    S = oracle.score_table(z)
    eng.set_scoring(S)
    eng.load_db(packed, clen)
    if tasks is None:
        tq, tc = all_vs_all(nq, len(clen))
        gt = None
    else:
        tq, tc = tasks
        gt = np.stack([tq, tc], 1).astype(np.uint32)
    ohits, obest = oracle.run_tasks(packed, off, clen, codes, qoff, budget, slot, nslots, tq, tc, S, mode, best=best)
    for seed in (True, False):          # both filters: pigeonhole seeds (where the batch allows) and the Myers prefix filter
        eng.set_seed_filter(seed)
        hits, gbest = eng.align(codes, qoff, budget, gt, mode, slot=slot, nslots=nslots, best=best)
        assert np.array_equal(gbest, obest), "per-slot minima differ (seed filter %s)" % seed
        assert len(hits) == len(ohits), (len(hits), len(ohits), seed)
        assert np.array_equal(hits, ohits), "hits differ (seed filter %s): first diff %s" % (seed,
            next(((tuple(a), tuple(b)) for a, b in zip(hits, ohits) if tuple(a) != tuple(b)), None),)
    eng.set_seed_filter(True)
    hits, gbest = eng.align(codes, qoff, budget, gt, mode, slot=slot, nslots=nslots, best=best)
    return hits, eng.stats()


@pytest.mark.parametrize("mode", [0, 1])
def test_c1_shape_all_vs_all(eng, oracle, mode):
    """configs[0] shape, scaled: 100 bp reads, 0-3 edits vs 1 kb references, -i 0.97 (budget 3)."""
    rng = np.random.default_rng(20261017)
    refs = synth.random_refs(96, 1000, rng)
    packed, off, clen = synth.pack_clumps(refs)
    reads, _ = synth.reads_from_clumps(packed, off, clen, 300, 100, 3, rng)
    budget = [oracle.budget(0.97, len(r)) for r in reads]
    hits, st = check(eng, oracle, packed, off, clen, reads, budget, mode=mode)
    assert len(hits) >= 290


def test_acx_shape_task_list_with_shared_slots(eng, oracle):
    """configs[1] shape: ~214-column clumps, 100 bp reads with exactly 2 edits, forward + reverse
    complement copies sharing one running minimum (burst.c:4218), explicit candidate lists."""
    rng = np.random.default_rng(5)
    refs = synth.random_refs(16 * 200, 214, rng, jitter=8)
    packed, off, clen = synth.pack_clumps(refs)
    reads, origin = synth.reads_from_clumps(packed, off, clen, 400, 100, 2, rng, exact_edits=True, rc_rate=0.5)
    fwd_rc = []
    for r in reads:
        fwd_rc += [r, synth.RC_TABLE[r[::-1]]]
    nq = len(fwd_rc)
    slot = (np.arange(nq) // 2).astype(np.uint32)
    tq, tc = [], []
    for i in range(nq):
        cands = {int(origin[i // 2, 0])} | set(int(v) for v in rng.integers(0, len(clen), 6))
        for c in sorted(cands):
            tq.append(i); tc.append(c)
    budget = [oracle.budget(0.98, len(r)) for r in fwd_rc]
    hits, st = check(eng, oracle, packed, off, clen, fwd_rc, budget,
                     tasks=(np.array(tq, np.uint32), np.array(tc, np.uint32)), slot=slot, nslots=nq // 2)
    assert len(hits) >= 380


@pytest.mark.parametrize("z", [1, 0])
def test_iupac_and_ragged(eng, oracle, z):
    """configs[4] shape: 50-320 bp queries with ambiguous bases, ragged clumps with pads, -i 0.95."""
    rng = np.random.default_rng(99 + z)
    refs = synth.random_refs(16 * 12, 500, rng, jitter=150, iupac_rate=0.01)
    packed, off, clen = synth.pack_clumps(refs)
    reads = []
    for i in range(120):
        L = int(rng.integers(50, 321))
        r, _ = synth.reads_from_clumps(packed, off, clen, 1, min(L, 340), int(0.04 * L), rng)
        r = r[0]
        m = rng.random(len(r)) < 0.005
        r[m] = rng.integers(5, 16, size=int(m.sum()), dtype=np.uint8)
        reads.append(r)
    budget = [oracle.budget(0.95, len(r)) for r in reads]
    hits, st = check(eng, oracle, packed, off, clen, reads, budget, mode=0, z=z)
    assert len(hits) > 50
    check(eng, oracle, packed, off, clen, reads[:40], budget[:40], mode=1, z=z)


def test_short_queries_and_edges(eng, oracle):
    """Queries shorter than the 32-row filter, hits hanging over either end of the window,
    a clump shorter than the query, zero budget, a query containing a non-letter (code 0)."""
    rng = np.random.default_rng(3)
    refs = synth.random_refs(31, 120, rng, jitter=30) + [rng.integers(1, 5, 40, dtype=np.uint8)]
    packed, off, clen = synth.pack_clumps(refs)
    reads = []
    for i in range(60):
        r = refs[int(rng.integers(0, len(refs)))]
        n = int(rng.integers(4, 60))
        kind = i % 4
        if kind == 0:   # inside
            o = int(rng.integers(0, max(1, len(r) - n))); q = r[o:o + n].copy()
        elif kind == 1:  # hangs over the left end by up to 2 bases
            h = int(rng.integers(1, 3)); q = np.concatenate([rng.integers(1, 5, h, dtype=np.uint8), r[:n]])
        elif kind == 2:  # hangs over the right end
            h = int(rng.integers(1, 3)); q = np.concatenate([r[-n:], rng.integers(1, 5, h, dtype=np.uint8)])
        else:
            q = synth.mutate(r[:n], 1, rng)
        reads.append(q.astype(np.uint8))
    bad = reads[5].copy(); bad[len(bad) // 2] = 0; reads.append(bad)
    reads.append(refs[0][:100].copy() if len(refs[0]) >= 100 else refs[0].copy())
    for budget in (0, 1, 2, 4):
        check(eng, oracle, packed, off, clen, reads, [budget] * len(reads), mode=0)
    check(eng, oracle, packed, off, clen, reads, [3] * len(reads), mode=1)


def test_large_budget_wide_bands(eng, oracle):
    """Budgets where the 32-row filter is weak: wide hulls, the 64-wide and the generic band kernels."""
    rng = np.random.default_rng(11)
    refs = synth.random_refs(32, 420, rng, jitter=20)
    packed, off, clen = synth.pack_clumps(refs)
    reads, _ = synth.reads_from_clumps(packed, off, clen, 24, 300, 20, rng)
    for k in (9, 16, 30):
        check(eng, oracle, packed, off, clen, reads, [k] * len(reads), mode=0)
    check(eng, oracle, packed, off, clen, reads[:8], [16] * 8, mode=1)


def test_running_minimum_carried_between_batches(eng, oracle):
    rng = np.random.default_rng(21)
    refs = synth.random_refs(64, 300, rng)
    packed, off, clen = synth.pack_clumps(refs)
    reads, _ = synth.reads_from_clumps(packed, off, clen, 50, 100, 3, rng)
    best = rng.integers(0, 4, len(reads)).astype(np.uint16)
    check(eng, oracle, packed, off, clen, reads, [3] * len(reads), mode=0, best=best)


def test_reference_shard_skips_foreign_clumps(eng, oracle):
    """first_clump != 0: tasks naming clumps of other shards are skipped."""
    rng = np.random.default_rng(8)
    refs = synth.random_refs(16 * 8, 250, rng)
    packed, off, clen = synth.pack_clumps(refs)
    reads, _ = synth.reads_from_clumps(packed, off, clen, 64, 100, 2, rng)
    codes, qoff = synth.concat_queries(reads)
    nq = len(reads); budget = np.full(nq, 2, np.uint16)
    tq, tc = all_vs_all(nq, len(clen))
    eng.set_scoring(oracle.score_table(1))
    full, fbest = None, None
    eng.load_db(packed, clen)
    full, fbest = eng.align(codes, qoff, budget, np.stack([tq, tc], 1))
    # two shards of 4 clumps each, minima combined by MIN, then select
    parts = []
    bests = []
    for s in range(2):
        lo, hi = s * 4, s * 4 + 4
        eng.load_db(packed[int(off[lo]):int(off[hi]) if hi < len(off) else len(packed)], clen[lo:hi], first_clump=lo)
        eng.upload(codes, qoff, budget, np.stack([tq, tc], 1))
        eng.run_extend(0)
        eng.run_select(1)      # keep everything within budget; the host applies the global minimum
        h, b = eng.download()
        parts.append(h); bests.append(b)
    gbest = np.minimum(bests[0], bests[1])
    assert np.array_equal(gbest, fbest)
    allh = np.concatenate(parts)
    keep = allh[allh["ed"] == gbest[tq[allh["task"]]]]
    keep = keep[np.lexsort((keep["lane"], keep["task"]))]
    assert np.array_equal(keep, full)


def test_run_lists_match_task_lists(eng, oracle):
    """bg_run work lists (one entry per clump visit of a bunch, burst.c:4137-4157) against the same
    visits written out as tasks for the oracle; hits come back keyed run * 16 + query-in-run."""
    from burst_b200.engine import RUN_DTYPE, RUN_MAX
    rng = np.random.default_rng(17)
    refs = synth.random_refs(16 * 64, 214, rng, jitter=6)
    packed, off, clen = synth.pack_clumps(refs)
    reads, origin = synth.reads_from_clumps(packed, off, clen, 203, 100, 2, rng, exact_edits=True)
    codes, qoff = synth.concat_queries(reads)
    nq = len(reads)
    budget = np.full(nq, 2, np.uint16)
    runs, tq, tc, key = [], [], [], []
    for q0 in range(0, nq, 11):                      # bunches of 11 (last one ragged)
        n = min(11, nq - q0)
        cands = sorted({int(origin[q, 0]) for q in range(q0, q0 + n)} | {int(v) for v in rng.integers(0, len(clen), 2)})
        for c in cands:
            for i in range(n):
                tq.append(q0 + i); tc.append(c); key.append(len(runs) * RUN_MAX + i)
            runs.append((c, q0, n))
    runs = np.array(runs, dtype=RUN_DTYPE)
    S = oracle.score_table(1)
    eng.set_scoring(S); eng.load_db(packed, clen)
    slot = np.arange(nq, dtype=np.uint32)
    ohits, obest = oracle.run_tasks(packed, off, clen, codes, qoff, budget, slot, nq, np.array(tq, np.uint32), np.array(tc, np.uint32), S, 0)
    ohits = ohits.copy(); ohits["task"] = np.array(key, np.uint32)[ohits["task"]]
    for seed in (True, False):
        eng.set_seed_filter(seed)
        hits, best = eng.align(codes, qoff, budget, None, 0, runs=runs)
        assert np.array_equal(best, obest)
        assert np.array_equal(hits, ohits), seed
    eng.set_seed_filter(True)
    eng.align(codes, qoff, budget, None, 0, runs=runs)
    st = eng.stats()
    assert st["seed_queries"] == nq and st["seed_stride"] == 8 and st["seed_window"] == 16
    assert len(hits) >= 200


def test_pipelined_one_call_path(eng, oracle):
    """bg_align_runs / bg_align_runs_into on a list long enough to be cut into slices (copy of slice i+1 overlapping
    the kernels of slice i): same hits and minima as the oracle, with reverse-complement-style shared slots whose
    two strands fall into different slices, running minima carried in, both selection modes, ragged bunches."""
    from burst_b200.engine import RUN_DTYPE, RUN_MAX, HIT_DTYPE, PARAM_PIPE_MIN_RUNS, PARAM_PIPE_SLICES
    rng = np.random.default_rng(41)
    refs = synth.random_refs(16 * 40, 214, rng, jitter=6)
    packed, off, clen = synth.pack_clumps(refs)
    reads, origin = synth.reads_from_clumps(packed, off, clen, 360, 100, 3, rng)
    # mix in a few queries the seed filter cannot take (IUPAC base, large budget)
    for i in range(0, len(reads), 37):
        reads[i][50] = 5
    codes, qoff = synth.concat_queries(reads)
    nq = len(reads)
    budget = np.full(nq, 3, np.uint16); budget[::53] = 9
    slot = (np.arange(nq, dtype=np.uint32) * 7) % (nq // 2)          # pairs of queries far apart share a slot
    nslots = nq // 2
    runs, tq, tc, key = [], [], [], []
    for q0 in range(0, nq, 13):
        n = min(13, nq - q0)
        cands = sorted({int(origin[q, 0]) for q in range(q0, q0 + n)} | {int(v) for v in rng.integers(0, len(clen), 2)})
        for c in cands:
            for i in range(n):
                tq.append(q0 + i); tc.append(c); key.append(len(runs) * RUN_MAX + i)
            runs.append((c, q0, n))
    runs = np.array(runs, dtype=RUN_DTYPE)
    S = oracle.score_table(1)
    eng.set_scoring(S); eng.load_db(packed, clen)
    best_in = np.full(nslots, 0xFFFF, np.uint16); best_in[::5] = 1
    try:
        eng.set_param(PARAM_PIPE_MIN_RUNS, 16)
        for mode in (0, 1):
            for bi in (None, best_in):
                ohits, obest = oracle.run_tasks(packed, off, clen, codes, qoff, budget, slot, nslots, np.array(tq, np.uint32), np.array(tc, np.uint32), S, mode, best=bi)
                ohits = ohits.copy(); ohits["task"] = np.array(key, np.uint32)[ohits["task"]]
                ohits = ohits[np.lexsort((ohits["lane"], ohits["task"]))]
                for slices in (8, 3, 0):
                    eng.set_param(PARAM_PIPE_SLICES, slices)
                    hits, best = eng.align(codes, qoff, budget, None, mode, slot=slot, nslots=nslots, best=bi, runs=runs)
                    assert np.array_equal(best, obest), (mode, slices)
                    assert np.array_equal(hits, ohits), (mode, slices)
                    buf = np.zeros(len(ohits) + 5, HIT_DTYPE); b2 = np.full(nslots, 0xFFFF, np.uint16) if bi is None else bi.copy()
                    n = eng.align_runs_into(codes, qoff, budget, runs, buf, b2, mode, slot=slot, nslots=nslots)
                    assert n == len(ohits) and np.array_equal(buf[:n], ohits) and np.array_equal(b2, obest)
        # a list in no particular order: every slice then spans (almost) the whole query array -- slower, same answers
        shuf = runs[rng.permutation(len(runs))]
        eng.set_param(PARAM_PIPE_SLICES, 0)
        want_h, want_b = eng.align(codes, qoff, budget, None, 0, slot=slot, nslots=nslots, runs=shuf)
        eng.set_param(PARAM_PIPE_SLICES, 4)
        got_h, got_b = eng.align(codes, qoff, budget, None, 0, slot=slot, nslots=nslots, runs=shuf)
        assert np.array_equal(got_h, want_h) and np.array_equal(got_b, want_b) and len(got_h) > 100
        with pytest.raises(RuntimeError, match="room for"):
            eng.align_runs_into(codes, qoff, budget, runs, np.zeros(3, HIT_DTYPE), None, 0, slot=slot, nslots=nslots)
        bad = runs.copy(); bad["query0"][len(bad) // 2] = nq - 2; bad["nq"][len(bad) // 2] = 5
        with pytest.raises(RuntimeError, match="malformed"):
            eng.align(codes, qoff, budget, None, 0, slot=slot, nslots=nslots, runs=bad)
    finally:
        eng.set_param(PARAM_PIPE_MIN_RUNS, 4096); eng.set_param(PARAM_PIPE_SLICES, 4)


def test_packed4_queries(eng, oracle):
    """BG_Q_PACKED4: the same batch given as a nibble stream (odd-length queries, so later ones start on a high nibble),
    through the resident path, the one-call path and its pipelined form; all-vs-all and run lists."""
    from burst_b200.engine import Engine, RUN_DTYPE, RUN_MAX, PARAM_PIPE_MIN_RUNS
    rng = np.random.default_rng(43)
    refs = synth.random_refs(16 * 10, 230, rng, jitter=30, iupac_rate=0.003)
    packed, off, clen = synth.pack_clumps(refs)
    reads = []
    for i in range(120):
        r, _ = synth.reads_from_clumps(packed, off, clen, 1, int(rng.integers(60, 131)), 3, rng)
        reads.append(r[0])
    codes, qoff = synth.concat_queries(reads)
    assert any(int(o) & 1 for o in qoff[1:-1])
    nq = len(reads)
    budget = np.full(nq, 3, np.uint16)
    p4 = ("packed4", Engine.pack4(codes))
    S = oracle.score_table(1)
    eng.set_scoring(S); eng.load_db(packed, clen)
    slot = np.arange(nq, dtype=np.uint32)
    for mode in (0, 1):
        want_h, want_b = eng.align(codes, qoff, budget, None, mode)                       # byte form, checked against the oracle elsewhere
        tq = np.tile(np.arange(nq, dtype=np.uint32), len(clen)); tc = np.repeat(np.arange(len(clen), dtype=np.uint32), nq)
        oh, ob = oracle.run_tasks(packed, off, clen, codes, qoff, budget, slot, nq, tq, tc, S, mode)
        assert np.array_equal(want_h, oh) and np.array_equal(want_b, ob)
        h, b = eng.align(p4, qoff, budget, None, mode)
        assert np.array_equal(h, want_h) and np.array_equal(b, want_b)
        eng.upload(p4, qoff, budget, None); eng.run(mode); h, b = eng.download()
        assert np.array_equal(h, want_h) and np.array_equal(b, want_b)
    runs = np.array([(c, q0, min(16, nq - q0)) for q0 in range(0, nq, 16) for c in range(len(clen))], dtype=RUN_DTYPE)
    want_h, want_b = eng.align(codes, qoff, budget, None, 0, runs=runs)
    try:
        for min_runs in (4096, 8):                                                        # single batch, then slices
            eng.set_param(PARAM_PIPE_MIN_RUNS, min_runs)
            h, b = eng.align(p4, qoff, budget, None, 0, runs=runs)
            assert np.array_equal(h, want_h) and np.array_equal(b, want_b)
    finally:
        eng.set_param(PARAM_PIPE_MIN_RUNS, 4096)
    assert len(want_h) >= 100


def test_repeats_give_several_seed_clusters(eng, oracle):
    """A read that occurs more than once in one reference lane (tandem / distant repeats, with and without
    errors) yields several diagonal clusters for one (query, lane); the merged result must be what the
    reference's left-to-right last-row scan gives (burst.c:826-883): best (score, shift), numGapR of its
    first occurrence, column of its last."""
    rng = np.random.default_rng(23)
    reads, refs = [], []
    for i in range(48):
        q = rng.integers(1, 5, 100, dtype=np.uint8)
        copies = [q.copy() for _ in range(int(rng.integers(2, 5)))]
        for j, cp in enumerate(copies):
            ne = int(rng.integers(0, 4))
            if (i + j) % 3 == 0:
                ne = 0
            copies[j] = synth.mutate(cp, ne, rng)
        parts = []
        for cp in copies:
            parts += [rng.integers(1, 5, int(rng.integers(3, 90)), dtype=np.uint8), cp]
        parts.append(rng.integers(1, 5, 20, dtype=np.uint8))
        refs.append(np.concatenate(parts).astype(np.uint8))
        reads.append(q)
    refs += synth.random_refs(16, 400, rng)
    packed, off, clen = synth.pack_clumps(refs)
    for k in (1, 2, 3):
        hits, st = check(eng, oracle, packed, off, clen, reads, [k] * len(reads), mode=0)
        assert len(hits) >= 20
    check(eng, oracle, packed, off, clen, reads, [3] * len(reads), mode=1)


def test_mixed_budgets_split_between_filters(eng, oracle):
    """One batch where some queries fit the seed automaton and others (large budget, short pieces) do not."""
    rng = np.random.default_rng(29)
    refs = synth.random_refs(16 * 6, 360, rng, jitter=25)
    packed, off, clen = synth.pack_clumps(refs)
    reads, budget = [], []
    for i in range(90):
        L = int(rng.choice([40, 100, 150, 260]))
        r, _ = synth.reads_from_clumps(packed, off, clen, 1, L, int(rng.integers(0, 4)), rng)
        reads.append(r[0]); budget.append(int(rng.choice([0, 1, 2, 3, 3, 3, 8, 13])))
    hits, st = check(eng, oracle, packed, off, clen, reads, budget, mode=0)
    assert 0 < st["seed_queries"] < len(reads)
    check(eng, oracle, packed, off, clen, reads, budget, mode=1)


def test_seed_layouts_and_tuning_knobs(eng, oracle):
    """The seed filter's window layouts -- a probe every 8 columns with 16-base windows (100 bp, budget 2),
    every 4 columns with shorter windows (budget 5: stretches of 16 bases) -- under window filters small
    enough to be mostly false positives and odd numbers of runs per warp.  Results never depend on the knobs."""
    from burst_b200.engine import PARAM_SEED_CHUNK, PARAM_SEED_WORDS, PARAM_SEED_STAGE
    rng = np.random.default_rng(37)
    refs = synth.random_refs(16 * 9, 230, rng, jitter=40, iupac_rate=0.002)
    packed, off, clen = synth.pack_clumps(refs)
    try:
        for k, want in ((2, (8, 16)), (3, (8, 16)), (4, (4, 16)), (5, (4, 13)), (6, (4, 11))):
            reads, _ = synth.reads_from_clumps(packed, off, clen, 150, 100, k, rng)
            for chunk, words, stage in ((8, 0, 0), (1, 128, 1), (5, 256, 0), (64, 4096, 1)):
                eng.set_param(PARAM_SEED_CHUNK, chunk); eng.set_param(PARAM_SEED_WORDS, words); eng.set_param(PARAM_SEED_STAGE, stage)
                hits, st = check(eng, oracle, packed, off, clen, reads, [k] * len(reads), mode=0)
                assert (st["seed_stride"], st["seed_window"]) == want, (k, st)
                assert st["seed_queries"] >= 0.7 * len(reads) and len(hits) >= 100
        eng.set_param(PARAM_SEED_CHUNK, 3); eng.set_param(PARAM_SEED_WORDS, 128)
        check(eng, oracle, packed, off, clen, reads, [6] * len(reads), mode=1)
    finally:
        eng.set_param(PARAM_SEED_CHUNK, 8); eng.set_param(PARAM_SEED_WORDS, 0); eng.set_param(PARAM_SEED_STAGE, 0)


def test_malformed_input_is_an_error_not_a_crash(eng, oracle):
    from burst_b200.engine import RUN_DTYPE
    rng = np.random.default_rng(31)
    refs = synth.random_refs(16, 200, rng)
    packed, off, clen = synth.pack_clumps(refs)
    eng.load_db(packed, clen)
    reads, _ = synth.reads_from_clumps(packed, off, clen, 8, 50, 1, rng)
    codes, qoff = synth.concat_queries(reads)
    with pytest.raises(RuntimeError, match="malformed"):
        eng.align(codes, qoff, np.full(8, 255, np.uint16), None)          # budget above 254
    with pytest.raises(RuntimeError, match="malformed"):
        eng.align(codes, qoff, np.full(8, 1, np.uint16), None, runs=np.array([(0, 4, 9)], dtype=RUN_DTYPE))   # runs past the batch
    with pytest.raises(RuntimeError, match="names query"):
        eng.align(codes, qoff, np.full(8, 1, np.uint16), np.array([[8, 0]], np.uint32))
    hits, best = eng.align(codes, qoff, np.full(8, 1, np.uint16), None)    # the context is still usable
    assert len(hits) >= 8


def test_bench_shape_at_scale(eng):
    """configs[1] shape at a size the oracle cannot check cell by cell (120 k reads = 240 k strands in 15 k bunches, 128 MB
    DB, ~120 k runs) through size-independent properties: every read is reported exactly at the lane it was cut from with
    at most its planted number of edits; the one-call path (pipelined, byte and nibble-packed input) returns byte-identical
    hits and minima to the resident path; both staging modes and the Myers filter agree."""
    from burst_b200.engine import Engine, RUN_DTYPE, HIT_DTYPE, PARAM_SEED_STAGE, PARAM_PIPE_MIN_RUNS
    w = synth.bunch_workload(120_000, 100, 2, 128 << 20, 214, seed=20261017)
    runs = np.ascontiguousarray(w["runs"], RUN_DTYPE)
    eng.set_scoring(default_scoring(1)); eng.load_db(w["packed"], w["clump_len"])
    eng.upload_runs(w["qcodes"], w["qoff"], w["budget"], runs, slot=w["slot"], nslots=w["nslots"])
    eng.run(0); hits, best = eng.download()
    assert int((best <= 2).sum()) == w["n_reads"]
    tq = runs["query0"][hits["task"] >> 4] + (hits["task"] & 15); tc = runs["clump"][hits["task"] >> 4]
    rd = w["slot"][tq]
    ok = (tc == w["true_clump"][rd]) & (hits["lane"] == w["true_lane"][rd]) & (hits["ed"] <= 2)
    assert len(np.unique(rd[ok])) == w["n_reads"]
    assert np.all(np.diff((hits["task"].astype(np.int64) << 4) | hits["lane"]) > 0)          # sorted by (task, lane), no duplicates
    buf = np.zeros(len(hits) + 16, HIT_DTYPE)
    try:
        for codes in (w["qcodes"], ("packed4", Engine.pack4(w["qcodes"]))):
            for stage in (1, 0):
                eng.set_param(PARAM_SEED_STAGE, stage)
                b2 = np.full(w["nslots"], 0xFFFF, np.uint16)
                n = eng.align_runs_into(codes, w["qoff"], w["budget"], runs, buf, b2, 0, slot=w["slot"], nslots=w["nslots"])
                assert n == len(hits) and np.array_equal(buf[:n], hits) and np.array_equal(b2, best)
        eng.set_seed_filter(False)                                                          # Myers prefix filter on a tenth of the list
        sub = runs[: len(runs) // 10]
        nqs = int(sub["query0"][-1] + sub["nq"][-1])
        h1, b1 = eng.align(w["qcodes"][: int(w["qoff"][nqs])], w["qoff"][: nqs + 1], w["budget"][:nqs], None, 0, slot=w["slot"][:nqs], nslots=w["nslots"], runs=sub)
        eng.set_seed_filter(True)
        h2, b2 = eng.align(w["qcodes"][: int(w["qoff"][nqs])], w["qoff"][: nqs + 1], w["budget"][:nqs], None, 0, slot=w["slot"][:nqs], nslots=w["nslots"], runs=sub)
        assert np.array_equal(h1, h2) and np.array_equal(b1, b2) and len(h1) > 1000
    finally:
        eng.set_seed_filter(True); eng.set_param(PARAM_SEED_STAGE, 0)


def test_seed_filter_with_tma_staging(eng, oracle):
    """BG_PARAM_SEED_STAGE = 1: clumps reach the seed filter through cp.async.bulk (TMA) + mbarrier staging, one run ahead,
    instead of the default direct 128-bit loads; 8, 4 and 2 groups per block.  Same results."""
    from burst_b200.engine import PARAM_SEED_STAGE, PARAM_SEED_GROUPS
    rng = np.random.default_rng(47)
    refs = synth.random_refs(16 * 9, 230, rng, jitter=40)
    packed, off, clen = synth.pack_clumps(refs)
    reads, _ = synth.reads_from_clumps(packed, off, clen, 150, 100, 2, rng)
    try:
        for stage, groups in ((1, 0), (1, 8), (0, 8), (1, 2), (0, 2)):
            eng.set_param(PARAM_SEED_STAGE, stage); eng.set_param(PARAM_SEED_GROUPS, groups)
            hits, st = check(eng, oracle, packed, off, clen, reads, [2] * len(reads), mode=0)
            assert len(hits) >= 100
    finally:
        eng.set_param(PARAM_SEED_STAGE, 0); eng.set_param(PARAM_SEED_GROUPS, 0)
