"""Multi-GPU drivers of the alignment path: one process per GPU, `torch.distributed` for the plumbing.

Two ways the path shards:

* QuerySharded -- every rank holds the whole DB and aligns a contiguous range of the batch's clump
  visits (runs, in bunch order).  No data-path collective.  Forward and reverse-complement strands
  of one read share a running minimum (burst.c:4218) and may land on different ranks, so the
  per-slot minima and hits are combined afterwards exactly like the reference combines its
  per-thread pods: global MIN over the minima, hits above the minimum dropped (burst.c:4497-4517).

* ReferenceSharded -- the DB does not fit one GPU: rank r holds clumps [lo_r, hi_r) and sees every
  query; after the filter + extend step ONE all-reduce(MIN) over the per-slot minima (u32 x nslots,
  in place on the engine's device array, NCCL over NVLink) tells every rank the global best distance,
  then each rank selects only its lanes at that minimum.  In FORAGE mode (all lanes within budget) no
  reduction is needed.

Both return, on every rank, this rank's hits; `gather_hits` assembles the union on all ranks for the
host-side reporters (CAPITALIST's reference counts are global, burst.c:4696-4727).
"""
import ctypes as C
import numpy as np
import torch
import torch.distributed as dist

from .engine import HIT_DTYPE, MODE_MIN, MODE_ALL, RUN_DTYPE, RUN_MAX


def split_range(n, world, rank):
    """Contiguous near-equal split of range(n)."""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def split_clumps(clump_len, world, rank):
    """Clump range of a reference shard: contiguous, balanced by packed bytes."""
    sizes = ((np.asarray(clump_len, np.uint64) + 1) // 2) * 16
    cum = np.concatenate([[0], np.cumsum(sizes)])
    tot = int(cum[-1])
    lo = int(np.searchsorted(cum, tot * rank // world, side="left"))
    hi = int(np.searchsorted(cum, tot * (rank + 1) // world, side="left")) if rank + 1 < world else len(clump_len)
    return min(lo, len(clump_len)), min(hi, len(clump_len))


class _DeviceView:
    """Zero-copy torch view of the engine's per-slot minima (u32 x nslots, values <= 0xFFFF)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3}


def _best_tensor(engine, nslots, on_cuda):
    ptr = engine.best_device_ptr()
    if on_cuda:
        return torch.as_tensor(_DeviceView(ptr, nslots), device="cuda")
    buf = (C.c_int32 * nslots).from_address(ptr)          # oracle-backed stand-in: host memory
    return torch.from_numpy(np.ctypeslib.as_array(buf))


def _all_gather_var(arr, group, device):
    """all_gather of a 1-d numpy structured/plain array of differing lengths -> list per rank."""
    world = dist.get_world_size(group)
    raw = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy()).to(device)
    n = torch.tensor([raw.numel()], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    pad = torch.zeros(cap, dtype=torch.uint8, device=device)
    pad[:raw.numel()] = raw
    outs = [torch.zeros(cap, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return [o[:s].cpu().numpy().view(arr.dtype) for o, s in zip(outs, sizes)]


class QuerySharded:
    def __init__(self, engine, group=None, on_cuda=True):
        self.eng, self.group, self.on_cuda = engine, group, on_cuda
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = "cuda" if on_cuda else "cpu"

    def load_db(self, packed, clump_len):
        self.eng.load_db(packed, clump_len)               # replicated

    def my_runs(self, runs):
        """This rank's contiguous share of the run list, cut at bunch boundaries (a change of query0)."""
        runs = np.ascontiguousarray(runs, RUN_DTYPE)
        n = len(runs)
        lo, hi = split_range(n, self.world, self.rank)

        def snap(i):
            while 0 < i < n and runs["query0"][i] == runs["query0"][i - 1]:
                i += 1
            return i
        return snap(lo), snap(hi)

    def align_runs(self, codes, offset, budget, runs, mode=MODE_MIN, slot=None, nslots=0, gather=True):
        """Returns (hits, best): with gather, the union over ranks (hit.task indexes the FULL run list) and
        the global minima; without, this rank's own."""
        runs = np.ascontiguousarray(runs, RUN_DTYPE)
        lo, hi = self.my_runs(runs)
        nq = len(offset) - 1
        if slot is None:
            slot = np.arange(nq, dtype=np.uint32); nslots = nq
        if hi > lo:
            hits, best = self.eng.align(codes, offset, budget, None, mode, slot=slot, nslots=nslots, runs=runs[lo:hi])
            hits = hits.copy(); hits["task"] += np.uint32(lo * RUN_MAX)
        else:
            hits, best = np.zeros(0, HIT_DTYPE), np.full(nslots, 0xFFFF, np.uint16)
        if not gather or self.world == 1:
            return hits, best
        b = torch.from_numpy(best.astype(np.int32)).to(self.device)
        dist.all_reduce(b, op=dist.ReduceOp.MIN, group=self.group)          # host-side merge of the minima (result assembly, not the DP path)
        gbest = b.cpu().numpy().astype(np.uint16)
        allh = np.concatenate(_all_gather_var(hits, self.group, self.device))
        if mode == MODE_MIN:                                                 # burst.c:4497-4517
            q = runs["query0"][allh["task"] // RUN_MAX] + allh["task"] % RUN_MAX
            allh = allh[allh["ed"] == gbest[np.asarray(slot)[q]]]
        allh = allh[np.lexsort((allh["lane"], allh["task"]))]
        return allh, gbest


class ReferenceSharded:
    """The engine must run on torch's current CUDA stream (Engine(device, stream=torch.cuda.current_stream().cuda_stream))
    so that the NCCL all-reduce is ordered between its extend and select steps without host synchronisation."""

    def __init__(self, engine, group=None, on_cuda=True):
        self.eng, self.group, self.on_cuda = engine, group, on_cuda
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = "cuda" if on_cuda else "cpu"
        self.lo = self.hi = 0

    def load_db(self, packed, clump_len):
        """Every rank passes the full host arrays (or at least its own slice's bytes); only this rank's
        clump range goes to its GPU."""
        clump_len = np.ascontiguousarray(clump_len, np.uint32)
        self.lo, self.hi = split_clumps(clump_len, self.world, self.rank)
        sizes = ((clump_len.astype(np.uint64) + 1) // 2) * 16
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
        if self.hi > self.lo:
            self.eng.load_db(packed[int(off[self.lo]):int(off[self.hi])], clump_len[self.lo:self.hi], first_clump=self.lo)
        return self.lo, self.hi

    def step_resident(self, mode, nslots, events=None):
        """One pass over the batch already uploaded to the engine: filter + extend on this rank's clump range, the all-reduce(MIN) of the
        per-slot minima in place on the device, selection against the combined minima -- all queued on the engine's stream (which must be
        torch's current stream), no host synchronisation in between.  `events` = (before, after) CUDA events recorded around the collective.
        Returns the device tensor of the combined minima (None on a rank without clumps, which still takes part in the collective)."""
        have = self.hi > self.lo
        if have:
            self.eng.run_extend(mode)                       # (settles a survivor-list overflow itself before it returns)
            best_t = _best_tensor(self.eng, nslots, self.on_cuda)
        else:
            best_t = torch.full((nslots,), 0xFFFF, dtype=torch.int32, device=self.device)
        if self.world > 1 and mode == MODE_MIN:
            if events:
                events[0].record(torch.cuda.current_stream())
            dist.all_reduce(best_t, op=dist.ReduceOp.MIN, group=self.group)  # the path's one collective (burst.c:4497-4517)
            if events:
                events[1].record(torch.cuda.current_stream())
        if have:
            self.eng.run_select(mode)
        return best_t

    def align_runs(self, codes, offset, budget, runs, mode=MODE_MIN, slot=None, nslots=0, gather=True):
        runs = np.ascontiguousarray(runs, RUN_DTYPE)
        nq = len(offset) - 1
        if slot is None:
            slot = np.arange(nq, dtype=np.uint32); nslots = nq
        have = self.hi > self.lo
        if have:
            self.eng.upload_runs(codes, offset, budget, runs, slot=slot, nslots=nslots)
        best_t = self.step_resident(mode, nslots)
        if have:
            hits, best = self.eng.download()                # (the first host synchronisation of the batch)
            if mode == MODE_MIN:
                best = best_t.cpu().numpy().astype(np.uint16)
        else:
            hits, best = np.zeros(0, HIT_DTYPE), best_t.cpu().numpy().astype(np.uint16)
        if not gather or self.world == 1:
            return hits, best
        if mode != MODE_MIN:
            b = torch.from_numpy(best.astype(np.int32)).to(self.device)
            dist.all_reduce(b, op=dist.ReduceOp.MIN, group=self.group)
            best = b.cpu().numpy().astype(np.uint16)
        allh = np.concatenate(_all_gather_var(hits, self.group, self.device))
        allh = allh[np.lexsort((allh["lane"], allh["task"]))]
        return allh, best
