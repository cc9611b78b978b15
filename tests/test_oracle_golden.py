"""CPU-only: the scalar oracle against committed golden vectors produced by the reference's own kernels
(scripts/make_golden.py -> tests/golden/kernels_z*.npz).  Runs without /root/reference."""
import os
import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(z):
    return np.load(os.path.join(GOLD, "kernels_z%d.npz" % z))


def cases(g):
    for i in range(len(g["clen"])):
        yield (i, g["packed"][g["packed_off"][i]:g["packed_off"][i + 1]], int(g["clen"][i]),
               g["q"][g["q_off"][i]:g["q_off"][i + 1]], int(g["emac"][i]))


@pytest.mark.parametrize("z", [0, 1])
def test_oracle_matches_golden_reference_vectors(oracle, z):
    g = load(z)
    S = oracle.score_table(z)
    nhit = 0
    for i, packed, clen, q, emac in cases(g):
        om, omins, ores = oracle.task(packed, clen, q, S, emac)
        rm = int(g["min"][i])
        if rm == 0xFFFFFFFF:
            assert om == 255
            continue
        assert np.array_equal(omins, g["mins"][i]), i
        if rm <= emac:
            nhit += 1
            for lane in range(16):
                if g["mins"][i][lane] > rm:
                    continue
                ed, gq, gr, fp = (int(v) for v in ores[lane])
                assert (ed, gq, gr, fp) == (int(g["mins"][i][lane]), int(g["gq"][i][lane]), int(g["gr"][i][lane]), int(g["fp"][i][lane]))
                assert oracle.identity(ed, len(q), gq) == g["score"][i][lane]
    assert nhit > 60
