"""Pin the scalar restatement (oracle/burst_oracle.c) against the UNMODIFIED reference kernels
(burst.c aded_mat16 / aded_mat16L / reScoreM_mat16 via oracle/ref_shim.c).  CPU only."""
import numpy as np
import pytest
from burst_b200 import synth


def test_tables_match_reference(oracle, reference):
    for z in (1, 0):
        reference.set_scoring(z)
        S, c2n, rvt = reference.tables()
        assert np.array_equal(S, oracle.score_table(z))          # burst.c:1310-1328
        assert np.array_equal(c2n, oracle.char2num())            # burst.c:1288-1307
        assert np.array_equal(rvt, oracle.rc_table())            # burst.c:168
    reference.set_scoring(1)


def test_budget_matches_reference(oracle, reference):
    for thres in (0.97, 0.98, 0.95, 0.9, 0.99, 0.5, 0.8):
        for n in list(range(1, 400)) + [1000, 6978, 20000]:
            assert oracle.budget(thres, n) == reference.budget(thres, n), (thres, n)
    assert oracle.budget(0.98, 147) == 2          # float32, not 3


def _case(rng, clen, qlen, edits, iupac_ref=0.0, iupac_q=0.0, ragged=False):
    refs = synth.random_refs(16, clen, rng, jitter=(clen // 4 if ragged else 0), iupac_rate=iupac_ref)
    packed, off, clens = synth.pack_clumps(refs)
    reads, origin = synth.reads_from_clumps(packed, off, clens, 1, qlen, edits, rng)
    q = reads[0]
    if iupac_q:
        m = rng.random(len(q)) < iupac_q
        q[m] = rng.integers(5, 16, size=int(m.sum()), dtype=np.uint8)
    return packed, int(clens[0]), q


@pytest.mark.parametrize("z", [1, 0])
@pytest.mark.parametrize("variant", [0, 1])
def test_two_passes_match_reference(oracle, reference, z, variant):
    rng = np.random.default_rng(1234 + z * 7 + variant)
    reference.set_scoring(z)
    S = oracle.score_table(z)
    n_hit = n_gap = 0
    for it in range(250):
        qlen = int(rng.integers(12, 140))
        clen = int(rng.integers(qlen + 5, 330))
        edits = int(rng.integers(0, 7))
        packed, clen, q = _case(rng, clen, qlen, edits,
                                iupac_ref=(0.02 if it % 3 == 0 else 0.0),
                                iupac_q=(0.03 if it % 4 == 0 else 0.0), ragged=(it % 2 == 0))
        emac = int(rng.integers(0, 9))
        rm, rmins, rscore, rfp, rgr, rgq = reference.task(packed, clen, q, emac, variant)
        om, omins, ores = oracle.task(packed, clen, q, S, emac)
        if rm == 0xFFFFFFFF:
            assert om == 255 and np.all(omins == 255)
            continue
        assert np.array_equal(rmins, omins), (it, rmins, omins)
        assert om == rm or (om == 255 and rm > emac)
        if rm <= emac:
            n_hit += 1
            for lane in range(16):
                if rmins[lane] > rm:
                    continue
                ed, gq, gr, fp = (int(v) for v in ores[lane])
                assert (ed, gq, gr, fp) == (int(rmins[lane]), int(rgq[lane]), int(rgr[lane]), int(rfp[lane])), (it, lane)
                assert oracle.identity(ed, len(q), gq) == rscore[lane]
                n_gap += (gq + gr) > 0
    reference.set_scoring(1)
    assert n_hit > 100 and n_gap > 20


def test_forage_bound_matches_reference(oracle, reference):
    """FORAGE/ANY hand pass 2 the budget instead of the minimum (burst.c:4224): lanes within
    budget must still get the same triple."""
    rng = np.random.default_rng(77)
    S = oracle.score_table(1)
    checked = 0
    for it in range(120):
        packed, clen, q = _case(rng, 260, 100, int(rng.integers(0, 5)))
        emac = 6
        rm, rmins, rscore, rfp, rgr, rgq = reference.task(packed, clen, q, emac, 1, rescore_ed=emac)
        om, omins, ores = oracle.task(packed, clen, q, S, emac, rescore_ed=emac)
        assert np.array_equal(rmins if rm != 0xFFFFFFFF else np.full(16, 255), omins)
        if rm <= emac:
            for lane in range(16):
                if rmins[lane] <= emac:
                    checked += 1
                    assert tuple(int(v) for v in ores[lane]) == (int(rmins[lane]), int(rgq[lane]), int(rgr[lane]), int(rfp[lane]))
    assert checked > 50
