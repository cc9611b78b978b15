"""CPU-only: the C-ABI library builds for sm_100a, loads, and exports every symbol the header declares."""
import ctypes
import os
import re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "burst_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from burst_b200 import engine
    if not os.path.exists(engine.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(engine.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libburst_b200.so does not export %s" % n
    assert set(engine.EXPORTS) == set(names)


def test_init_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from burst_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CUDA device|CPU fallback"):
        Engine(0)


def test_default_scoring_is_the_reference_table(oracle):
    import numpy as np
    from burst_b200 import engine
    for z in (0, 1):
        assert np.array_equal(engine.default_scoring(z), oracle.score_table(z))
