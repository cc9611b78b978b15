"""GPU parity at the shapes of BASELINE.json configs[2] and configs[3] (scaled so that the scalar oracle finishes in
seconds), through the C ABI, for both seed-filter forms and the Myers prefix filter.

  C3  amplicon: 292 bp reads, -i 0.97 (budget 9), references in a mutation tree so that the lanes of a clump are
      near-identical (most visited lanes seed, many tie), sorted strands in bunches of 16, >= 30 clump visits per
      query, MIN and ALL selection.                                         burst.c:4136-4284, BASELINE.md section 2
  C4  150 bp reads, -i 0.98 (budget 3), the DB cut into two reference shards whose per-slot minima are combined by
      MIN before the selection (the rule of burst.c:4490-4519).
"""
import numpy as np
import pytest
from burst_b200 import synth
from burst_b200.engine import RUN_DTYPE, PARAM_SEED_IMPL, PARAM_SEED_NCH, PARAM_SEED_FB

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from burst_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def amplicon_case(seed, n_refs=16 * 36, n_reads=48, read_len=292, ref_len=700):
    rng = np.random.default_rng(seed)
    refs = synth.mutation_tree_refs(rng, n_refs, ref_len, levels=(0.25, 0.10, 0.05, 0.006))   # 9 leaves per genus, a few edits apart
    # the reference sorts references so that similar ones share clumps; a lexicographic sort does that for a mutation tree
    refs.sort(key=lambda r: r.tobytes())
    packed, off, clen = synth.pack_clumps(refs)
    reads, src = synth.amplicon_reads(refs, n_reads, read_len, 5, rng, start=60, jitter=5, dup_rate=0.0)
    # forward + reverse complement strands, sorted (burst.c:3087-3109, 3181-3184)
    strands = []; sread = []
    for i, r in enumerate(reads):
        strands += [r, synth.RC_TABLE[r[::-1]]]; sread += [i, i]
    codes, qoff = synth.concat_queries(strands)
    order = synth.sort_strands(codes, qoff)
    strands = [strands[i] for i in order]; slot = np.array([sread[i] for i in order], np.uint32)
    return rng, refs, packed, off, clen, strands, slot, len(reads)


@pytest.mark.parametrize("mode", [0, 1])
def test_c3_amplicon_shape(eng, oracle, mode):
    rng, refs, packed, off, clen, strands, slot, nslots = amplicon_case(2026 + mode)
    codes, qoff = synth.concat_queries(strands)
    nq = len(strands)
    budget = np.array([oracle.budget(0.97, len(s)) for s in strands], np.uint16)
    assert int(budget.max()) == 9
    nclumps = len(clen)
    # every bunch visits 32 clumps: a window of the sorted DB around a random point (near-identical families are adjacent)
    def cands(b, q0, n):
        c0 = int(rng.integers(0, nclumps - 32))
        v = np.arange(c0, c0 + 32); rng.shuffle(v)
        return v
    runs, tq, tc, key = synth.bunch_runs(nq, 16, cands)
    S = oracle.score_table(1)
    ohits, obest = oracle.run_tasks(packed, off, clen, codes, qoff, budget, slot, nslots, tq, tc, S, mode)
    ohits = ohits.copy(); ohits["task"] = key[ohits["task"]]
    ohits = ohits[np.lexsort((ohits["lane"], ohits["task"]))]
    assert len(ohits) > (60 if mode == 0 else 200)                        # families: many lanes within budget / tied
    eng.set_scoring(S); eng.load_db(packed, clen)
    try:
        for impl, nch, fb, seedf in ((1, 8, 1, True), (1, 4, 2, True), (1, 8, 2, True), (0, 8, 1, True), (1, 8, 1, False)):
            eng.set_param(PARAM_SEED_IMPL, impl); eng.set_param(PARAM_SEED_NCH, nch); eng.set_param(PARAM_SEED_FB, fb); eng.set_seed_filter(seedf)
            hits, best = eng.align(codes, qoff, budget, None, mode, slot=slot, nslots=nslots, runs=runs.astype(RUN_DTYPE))
            assert np.array_equal(best, obest), (impl, nch, fb, seedf)
            assert len(hits) == len(ohits) and np.array_equal(hits, ohits), (impl, nch, fb, seedf, len(hits), len(ohits))
    finally:
        eng.set_param(PARAM_SEED_IMPL, 1); eng.set_param(PARAM_SEED_NCH, 8); eng.set_param(PARAM_SEED_FB, 1); eng.set_seed_filter(True)
    eng.align(codes, qoff, budget, None, mode, slot=slot, nslots=nslots, runs=runs.astype(RUN_DTYPE))
    st = eng.stats()
    assert st["seed_queries"] == nq and st["seed_stride"] == 8


@pytest.mark.parametrize("tiny", [False, True])
def test_c4_shape_two_reference_shards(eng, oracle, tiny):
    """tiny: each shard starts from a 32-entry survivor list, so bg_batch_run_extend has to notice the overflow, grow the list and redo
    filter + extend BEFORE the minima are handed to the MIN-combination (a truncated list must never reach the all-reduce)."""
    rng = np.random.default_rng(404)
    refs = synth.random_refs(16 * 48, 306, rng, jitter=4)
    # a few references repeated in the other half of the DB, so that both shards hold hits of one read
    for i in range(0, 16 * 24, 29):
        refs[16 * 24 + i] = synth.mutate(refs[i], int(rng.integers(0, 3)), rng)
    packed, off, clen = synth.pack_clumps(refs)
    reads, origin = synth.reads_from_clumps(packed, off, clen, 220, 150, 3, rng, rc_rate=0.0)
    codes, qoff = synth.concat_queries(reads)
    nq = len(reads)
    budget = np.array([oracle.budget(0.98, len(r)) for r in reads], np.uint16)
    assert int(budget.max()) == 3
    nclumps = len(clen)
    def cands(b, q0, n):
        own = {int(origin[q, 0]) for q in range(q0, q0 + n)}
        twin = {(c + 24) % nclumps for c in own}
        return sorted(own | twin | {int(v) for v in rng.integers(0, nclumps, 3)})
    runs, tq, tc, key = synth.bunch_runs(nq, 16, cands)
    runs = runs.astype(RUN_DTYPE)
    S = oracle.score_table(1)
    slot = np.arange(nq, dtype=np.uint32)
    ohits, obest = oracle.run_tasks(packed, off, clen, codes, qoff, budget, slot, nq, tq, tc, S, 0)
    ohits = ohits.copy(); ohits["task"] = key[ohits["task"]]
    ohits = ohits[np.lexsort((ohits["lane"], ohits["task"]))]
    eng.set_scoring(S)
    eng.load_db(packed, clen)
    hits, best = eng.align(codes, qoff, budget, None, 0, runs=runs)
    assert np.array_equal(best, obest) and np.array_equal(hits, ohits)
    # two shards: extend on each, MIN of the minima, then select against the combined minima
    half = nclumps // 2
    parts, bests = [], []
    for lo, hi in ((0, half), (half, nclumps)):
        end = int(off[hi]) if hi < nclumps else len(packed)
        eng.load_db(packed[int(off[lo]):end], clen[lo:hi], first_clump=lo)
        if tiny:
            eng.set_surv_cap(32)
        eng.upload_runs(codes, qoff, budget, runs)
        eng.run_extend(0); eng.run_select(1)
        h, b = eng.download()
        parts.append(h); bests.append(b)
    gbest = np.minimum(bests[0], bests[1])
    assert np.array_equal(gbest, obest)
    allh = np.concatenate(parts)
    q_of = runs["query0"][allh["task"] >> 4] + (allh["task"] & 15)
    keep = allh[allh["ed"] == gbest[q_of]]
    keep = keep[np.lexsort((keep["lane"], keep["task"]))]
    assert np.array_equal(keep, ohits)
    assert len(set(np.nonzero(bests[0] != 0xFFFF)[0]) & set(np.nonzero(bests[1] != 0xFFFF)[0])) > 0   # some reads hit in both shards


@pytest.mark.parametrize("mode", [0, 1])
def test_compact_strand_batches_match_the_oracle(eng, oracle, mode):
    """bg_align_bunches_into (reads sent once at 2 / 4 bits per base, strands derived on the device, runs from the bunch lists)
    against the oracle on the written-out form; IUPAC reads included in the 4-bit case."""
    from burst_b200.engine import Engine, HIT_DTYPE
    rng = np.random.default_rng(77 + mode)
    refs = synth.random_refs(16 * 24, 214, rng, jitter=8)
    packed, off, clen = synth.pack_clumps(refs)
    reads, origin = synth.reads_from_clumps(packed, off, clen, 301, 100, 2, rng, rc_rate=0.5)
    reads = [r[:int(rng.integers(70, 101))] for r in reads]
    S = oracle.score_table(1)
    eng.set_scoring(S); eng.load_db(packed, clen)
    for packed2 in (True, False):
        if not packed2:
            for i in range(0, len(reads), 23):
                reads[i] = reads[i].copy(); reads[i][len(reads[i]) // 2] = 5              # an N: those strands go to the Myers filter
        budgets = [oracle.budget(0.98, len(r)) for r in reads]
        B = synth.strand_batch(reads, budgets, 16, lambda b, rd, rc: sorted({int(origin[r, 0]) for r in rd} | {int(v) for v in rng.integers(0, len(clen), 2)}))
        ohits, obest = oracle.run_tasks(packed, off, clen, B["qcodes"], B["qoff"], B["budget"], B["slot"], B["nreads"], B["tq"], B["tc"], S, mode)
        ohits = ohits.copy(); ohits["task"] = B["key"][ohits["task"]]
        ohits = ohits[np.lexsort((ohits["lane"], ohits["task"]))]
        stream = Engine.pack2(B["rcodes"]) if packed2 else Engine.pack4(B["rcodes"])
        buf = np.zeros(len(ohits) + 8, HIT_DTYPE); b2 = np.full(B["nreads"], 0xFFFF, np.uint16)
        n = eng.align_bunches_into(stream, B["rlen"], B["rbudget"], B["strand"], 16, B["cand_off"], B["cand"], buf, b2, mode, packed2=packed2)
        assert n == len(ohits) and np.array_equal(buf[:n], ohits), (mode, packed2, n, len(ohits))
        assert np.array_equal(b2, obest)
        assert n > 150
        # a survivor list that is too small: the call notices, enlarges it and redoes the batch (second attempt copies everything again)
        eng.set_surv_cap(32)
        buf[:] = 0; b2[:] = 0xFFFF
        n = eng.align_bunches_into(stream, B["rlen"], B["rbudget"], B["strand"], 16, B["cand_off"], B["cand"], buf, b2, mode, packed2=packed2)
        assert n == len(ohits) and np.array_equal(buf[:n], ohits) and np.array_equal(b2, obest), "after a survivor-list overflow"


@pytest.mark.parametrize("mode", [0, 1])
def test_long_queries_large_budgets(eng, oracle, mode):
    """The shape of the reference's manuscript data set (genes of up to 7 kbp against sheared genomes, budgets
    up to 139): queries of 150 - 1500 bases with budgets of 7 - 150 against 2600-column clumps that hold a repeat.  Queries with more
    than 16 stretches cannot take the pigeonhole filter: they go through the multi-word Myers prefix filter (k_filter<NW>, one
    instance per prefix length), every seed cluster its own survivor, bands wider than 64 in the per-thread global scratch."""
    rng = np.random.default_rng(900 + mode)
    refs = synth.random_refs(16 * 5, 2600, rng, jitter=30)
    for i in range(0, 80, 7):                                             # a repeat: the same 900 bases twice in one reference, 1100 apart
        refs[i][1500:2400] = refs[i][400:1300]
    for i in range(3, 80, 11):                                            # and a near copy of a reference in another clump
        refs[(i + 37) % 80] = synth.mutate(refs[i], 12, rng)
    packed, off, clen = synth.pack_clumps(refs)
    reads, budget = [], []
    for i in range(20):
        L = int(rng.choice([150, 300, 640, 900, 1500])); ident = float(rng.choice([0.95, 0.92, 0.90]))
        src = refs[int(rng.integers(0, 80))]; st = int(rng.integers(0, len(src) - L))
        k = oracle.budget(ident, L)
        r = synth.mutate(src[st:st + L], int(rng.integers(0, k + 1)), rng, p_sub=0.7, p_ins=0.15)
        if i % 6 == 5:
            r[len(r) // 3] = 5                                             # an N
        reads.append(r); budget.append(oracle.budget(ident, len(r)))
    budget = np.array(budget, np.uint16)
    assert int(budget.max()) > 100 and int(budget.min()) < 16
    codes, qoff = synth.concat_queries(reads)
    order = synth.sort_strands(codes, qoff)
    reads = [reads[i] for i in order]; budget = budget[order]
    codes, qoff = synth.concat_queries(reads)
    nq = len(reads)
    runs, tq, tc, key = synth.bunch_runs(nq, 16, lambda b, q0, n: np.arange(len(clen)))
    S = oracle.score_table(1)
    ohits, obest = oracle.run_tasks(packed, off, clen, codes, qoff, budget, np.arange(nq, dtype=np.uint32), nq, tq, tc, S, mode)
    ohits = ohits.copy(); ohits["task"] = key[ohits["task"]]
    ohits = ohits[np.lexsort((ohits["lane"], ohits["task"]))]
    assert len(ohits) >= nq
    eng.set_scoring(S); eng.load_db(packed, clen)
    try:
        for seedf in (True, False):
            eng.set_seed_filter(seedf)
            hits, best = eng.align(codes, qoff, budget, None, mode, runs=runs.astype(RUN_DTYPE))
            assert np.array_equal(best, obest), seedf
            assert len(hits) == len(ohits) and np.array_equal(hits, ohits), (seedf, len(hits), len(ohits))
    finally:
        eng.set_seed_filter(True)


def test_two_contexts_share_one_database(eng, oracle):
    """bg_share_db: two contexts on one GPU over one database in HBM, one host thread each, the one-call path at the same time
    (the bench's in-flight e2e leg): both must return what a single context returns."""
    import threading
    from burst_b200.engine import Engine, HIT_DTYPE
    rng = np.random.default_rng(77)
    refs = synth.random_refs(16 * 40, 230, rng, jitter=10)
    packed, off, clen = synth.pack_clumps(refs)
    reads, origin = synth.reads_from_clumps(packed, off, clen, 320, 100, 2, rng)
    codes, qoff = synth.concat_queries(reads)
    budget = np.full(len(reads), 2, np.uint16)
    runs = np.array([(int(origin[q, 0]), q, 1) for q in range(len(reads))] + [(c, q0, 16) for q0 in range(0, len(reads), 16) for c in (0, 7, len(clen) - 1)], dtype=RUN_DTYPE)
    eng.load_db(packed, clen)
    want_h, want_b = eng.align(codes, qoff, budget, None, 0, runs=runs)
    assert len(want_h) >= len(reads)
    other = Engine(0); other.share_db(eng)
    out = {}

    def work(name, e):
        for it in range(6):
            buf = np.zeros(len(want_h) + 8, HIT_DTYPE); best = np.full(len(reads), 0xFFFF, np.uint16)
            n = e.align_runs_into(codes, qoff, budget, runs, buf, best, 0)
            out[name, it] = (buf[:n].copy(), best)
    th = [threading.Thread(target=work, args=("a", eng)), threading.Thread(target=work, args=("b", other))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert len(out) == 12
    for h, b in out.values():
        assert np.array_equal(h, want_h) and np.array_equal(b, want_b)
    other.close()
    h, b = eng.align(codes, qoff, budget, None, 0, runs=runs)              # the owner still holds its database
    assert np.array_equal(h, want_h) and np.array_equal(b, want_b)


def test_compact_2bit_reads_of_any_length(eng, oracle):
    """k_compact_prep2 (one kernel from 2-bit packed reads to code bytes, nibble words and seed classes, both orientations): reads
    shorter than one 16-base step, longer than the 128 bases the eight threads of a strand cover per round, lengths that are not a
    multiple of 16 or 8, the stream's first read met through its reverse complement -- against the oracle on the written-out strands."""
    from burst_b200.engine import Engine, HIT_DTYPE
    rng = np.random.default_rng(4242)
    refs = synth.random_refs(16 * 12, 330, rng, jitter=6)
    packed, off, clen = synth.pack_clumps(refs)
    reads, origin = synth.reads_from_clumps(packed, off, clen, 90, 300, 4, rng, rc_rate=0.5)
    lens = [17, 300, 129, 128, 127, 16, 33, 47, 250, 18] + [int(v) for v in rng.integers(18, 301, len(reads) - 10)]
    reads = [r[:L] for r, L in zip(reads, lens)]
    S = oracle.score_table(1)
    eng.set_scoring(S); eng.load_db(packed, clen)
    budgets = [oracle.budget(0.97, len(r)) for r in reads]
    B = synth.strand_batch(reads, budgets, 16, lambda b, rd, rc: sorted({int(origin[r, 0]) for r in rd} | {int(v) for v in rng.integers(0, len(clen), 1)}))
    ohits, obest = oracle.run_tasks(packed, off, clen, B["qcodes"], B["qoff"], B["budget"], B["slot"], B["nreads"], B["tq"], B["tc"], S, 0)
    ohits = ohits.copy(); ohits["task"] = B["key"][ohits["task"]]
    ohits = ohits[np.lexsort((ohits["lane"], ohits["task"]))]
    buf = np.zeros(len(ohits) + 8, HIT_DTYPE); b2 = np.full(B["nreads"], 0xFFFF, np.uint16)
    n = eng.align_bunches_into(Engine.pack2(B["rcodes"]), B["rlen"], B["rbudget"], B["strand"], 16, B["cand_off"], B["cand"], buf, b2, 0, packed2=True)
    assert n == len(ohits) and np.array_equal(buf[:n], ohits), (n, len(ohits))
    assert np.array_equal(b2, obest)
    assert n >= 60
    # the same reads at 4 bits per base go through the two-kernel preparation: same answer
    buf4 = np.zeros(len(ohits) + 8, HIT_DTYPE); b4 = np.full(B["nreads"], 0xFFFF, np.uint16)
    n4 = eng.align_bunches_into(Engine.pack4(B["rcodes"]), B["rlen"], B["rbudget"], B["strand"], 16, B["cand_off"], B["cand"], buf4, b4, 0, packed2=False)
    assert n4 == n and np.array_equal(buf4[:n], buf[:n]) and np.array_equal(b4, b2)
