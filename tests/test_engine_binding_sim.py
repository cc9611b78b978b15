"""CPU-only: the ctypes mirror (burst_b200/engine.py) over the oracle-backed stand-in for the ABI (oracle/_sim) --
the input forms and entry points added for the pipelined one-call path must behave like the plain ones: nibble-packed
queries (BG_Q_PACKED4), caller-owned hit buffers (bg_align_runs_into), the overflow error.  The CUDA engine's versions
of the same checks are in tests/test_gpu_parity.py (test_packed4_queries, test_pipelined_one_call_path)."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMLIB = os.path.join(ROOT, "oracle", "_sim", "libburst_b200_sim.so")


@pytest.fixture(scope="module")
def sim():
    if not os.path.exists(SIMLIB):
        import __graft_entry__ as g
        g.build()
    from burst_b200.engine import Engine
    e = Engine(0, lib_path=SIMLIB)
    yield e
    e.close()


def small_batch(seed=5):
    from burst_b200 import synth
    from burst_b200.engine import RUN_DTYPE
    rng = np.random.default_rng(seed)
    refs = synth.random_refs(16 * 6, 220, rng, jitter=20)
    packed, off, clen = synth.pack_clumps(refs)
    reads = []
    origin = []
    for i in range(40):
        r, o = synth.reads_from_clumps(packed, off, clen, 1, int(rng.integers(61, 120)) | 1, 2, rng)   # odd lengths
        reads.append(r[0]); origin.append(int(o[0, 0]))
    codes, qoff = synth.concat_queries(reads)
    runs = np.array([(origin[q0], q0, min(8, len(reads) - q0)) for q0 in range(0, len(reads), 8)] +
                    [(c, q0, min(8, len(reads) - q0)) for q0 in range(0, len(reads), 8) for c in (0, len(clen) - 1)], dtype=RUN_DTYPE)
    return packed, clen, codes, qoff, np.full(len(reads), 2, np.uint16), runs


def test_packed4_and_caller_owned_buffers(sim):
    from burst_b200.engine import Engine, HIT_DTYPE
    packed, clen, codes, qoff, budget, runs = small_batch()
    sim.load_db(packed, clen)
    assert any(int(o) & 1 for o in qoff[1:-1])
    for mode in (0, 1):
        want_h, want_b = sim.align(codes, qoff, budget, None, mode, runs=runs)
        h, b = sim.align(("packed4", Engine.pack4(codes)), qoff, budget, None, mode, runs=runs)
        assert np.array_equal(h, want_h) and np.array_equal(b, want_b)
        buf = np.zeros(len(want_h) + 3, HIT_DTYPE); b2 = np.full(len(budget), 0xFFFF, np.uint16)
        n = sim.align_runs_into(("packed4", Engine.pack4(codes)), qoff, budget, runs, buf, b2, mode)
        assert n == len(want_h) and np.array_equal(buf[:n], want_h) and np.array_equal(b2, want_b)
    assert len(want_h) > 0
    with pytest.raises(RuntimeError):
        sim.align_runs_into(codes, qoff, budget, runs, np.zeros(1, HIT_DTYPE), None, 1)


def test_pack4_layout():
    from burst_b200.engine import Engine
    c = np.array([1, 2, 3, 4, 15], np.uint8)
    assert Engine.pack4(c).tolist() == [0x21, 0x43, 0x0F]


def test_compact_strand_batches(sim):
    """bg_align_bunches_into: reads sent once (2 or 4 bits per base), strands + bunch lists -> the same hits and minima as the
    general form (strands as byte codes, runs written out)."""
    from burst_b200 import synth
    from burst_b200.engine import Engine, HIT_DTYPE
    rng = np.random.default_rng(12)
    refs = synth.random_refs(16 * 8, 230, rng, jitter=10)
    packed, off, clen = synth.pack_clumps(refs)
    reads, origin = synth.reads_from_clumps(packed, off, clen, 37, 100, 2, rng, rc_rate=0.5)
    reads = [r[:int(rng.integers(80, 101))] for r in reads]                       # ragged lengths
    B = synth.strand_batch(reads, [2] * len(reads), 8, lambda b, rd, rc: sorted({int(origin[r, 0]) for r in rd} | {0, len(clen) - 1}))
    sim.load_db(packed, clen)
    for mode in (0, 1):
        want_h, want_b = sim.align(B["qcodes"], B["qoff"], B["budget"], None, mode, slot=B["slot"], nslots=B["nreads"], runs=B["runs"])
        assert len(want_h) >= 30
        for packed2 in (False, True):
            stream = Engine.pack2(B["rcodes"]) if packed2 else Engine.pack4(B["rcodes"])
            buf = np.zeros(len(want_h) + 2, HIT_DTYPE); b2 = np.full(B["nreads"], 0xFFFF, np.uint16)
            n = sim.align_bunches_into(stream, B["rlen"], B["rbudget"], B["strand"], 8, B["cand_off"], B["cand"], buf, b2, mode, packed2=packed2)
            assert n == len(want_h) and np.array_equal(buf[:n], want_h) and np.array_equal(b2, want_b), (mode, packed2)


def test_share_db_second_context(sim):
    """bg_share_db: a second context that borrows the first one's database gives the same hits; it can be freed first, and loading
    a database of its own drops the borrowed one without touching the owner's."""
    from burst_b200.engine import Engine
    packed, clen, codes, qoff, budget, runs = small_batch(seed=9)
    sim.load_db(packed, clen)
    want_h, want_b = sim.align(codes, qoff, budget, None, 0, runs=runs)
    other = Engine(0, lib_path=SIMLIB)
    with pytest.raises(RuntimeError):
        Engine(0, lib_path=SIMLIB).share_db(other)          # nothing to share yet
    other.share_db(sim)
    h, b = other.align(codes, qoff, budget, None, 0, runs=runs)
    assert np.array_equal(h, want_h) and np.array_equal(b, want_b)
    other.load_db(packed[: int(((clen[:16].astype(np.uint64) + 1) // 2 * 16).sum())], clen[:16])     # its own, smaller database
    other.close()
    h, b = sim.align(codes, qoff, budget, None, 0, runs=runs)                                        # the owner's is intact
    assert np.array_equal(h, want_h) and np.array_equal(b, want_b)
