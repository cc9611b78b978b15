"""CPU-only, world_size 2 over gloo: the multi-GPU drivers' host logic (burst_b200/sharded.py) --
query sharding (no data-path collective, minima/hits merged afterwards) and reference sharding
(all-reduce(MIN) on the per-slot minima between extend and select) -- must give exactly the
single-process result.  The engine behind the ABI is the oracle-backed stand-in (oracle/_sim), so
this tier checks the sharding logic, not the kernels; tests/test_gpu_parity.py::test_reference_shard*
and scripts/gpu_multi_check.py cover the CUDA engine."""
import os
import subprocess
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMLIB = os.path.join(ROOT, "oracle", "_sim", "libburst_b200_sim.so")


def workload(seed=3):
    sys.path.insert(0, ROOT)
    from burst_b200 import synth
    from burst_b200.engine import RUN_DTYPE
    rng = np.random.default_rng(seed)
    refs = synth.random_refs(16 * 24, 214, rng, jitter=5)
    packed, off, clen = synth.pack_clumps(refs)
    reads, origin = synth.reads_from_clumps(packed, off, clen, 60, 100, 2, rng, exact_edits=True, rc_rate=0.5)
    strands = []
    for r in reads:
        strands += [r, synth.RC_TABLE[r[::-1]]]
    # sorted strands, as the reference's query preprocessing leaves them (burst.c:3021): fwd/rc of a read scatter
    order = sorted(range(len(strands)), key=lambda i: bytes(strands[i]))
    strands = [strands[i] for i in order]
    slot = np.array([i // 2 for i in order], np.uint32)
    codes, qoff = synth.concat_queries(strands)
    nq = len(strands)
    runs = []
    for q0 in range(0, nq, 16):
        n = min(16, nq - q0)
        cands = sorted({int(origin[int(slot[q]), 0]) for q in range(q0, q0 + n)})
        runs += [(c, q0, n) for c in cands]
    runs = np.array(runs, dtype=RUN_DTYPE)
    budget = np.full(nq, 2, np.uint16)
    return packed, clen, codes, qoff, budget, slot, len(reads), runs


def _worker(rank, world, port, kind, mode, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from burst_b200.engine import Engine
    from burst_b200 import sharded
    packed, clen, codes, qoff, budget, slot, nslots, runs = workload()
    eng = Engine(0, lib_path=SIMLIB)
    drv = (sharded.QuerySharded if kind == "queries" else sharded.ReferenceSharded)(eng, on_cuda=False)
    drv.load_db(packed, clen)
    hits, best = drv.align_runs(codes, qoff, budget, runs, mode, slot=slot, nslots=nslots)
    np.save(os.path.join(out, "hits_%d.npy" % rank), hits); np.save(os.path.join(out, "best_%d.npy" % rank), best)
    dist.destroy_process_group()


@pytest.fixture(scope="module")
def simlib():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "sim"], check=True)
    return SIMLIB


@pytest.mark.parametrize("kind", ["queries", "references"])
@pytest.mark.parametrize("mode", [0, 1])
def test_two_ranks_equal_one(simlib, tmp_path, kind, mode):
    sys.path.insert(0, ROOT)
    from burst_b200.engine import Engine
    packed, clen, codes, qoff, budget, slot, nslots, runs = workload()
    eng = Engine(0, lib_path=simlib)
    eng.load_db(packed, clen)
    want_hits, want_best = eng.align(codes, qoff, budget, None, mode, slot=slot, nslots=nslots, runs=runs)
    assert len(want_hits) >= 55
    port = 29500 + (os.getpid() + 7 * mode + (3 if kind == "queries" else 0)) % 2000
    mp.spawn(_worker, args=(2, port, kind, mode, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        hits = np.load(tmp_path / ("hits_%d.npy" % rank)); best = np.load(tmp_path / ("best_%d.npy" % rank))
        assert np.array_equal(best, want_best), (kind, mode, rank)
        assert np.array_equal(hits, want_hits), (kind, mode, rank)


def test_split_helpers():
    sys.path.insert(0, ROOT)
    from burst_b200 import sharded
    assert [sharded.split_range(10, 3, r) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]
    clen = np.array([100, 300, 100, 100, 200], np.uint32)
    parts = [sharded.split_clumps(clen, 2, r) for r in range(2)]
    assert parts[0][0] == 0 and parts[0][1] == parts[1][0] and parts[1][1] == 5
