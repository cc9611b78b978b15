"""ctypes binding of include/burst_b200.h (libburst_b200.so) for tests and bench.py.

This is a thin mirror of the C ABI -- the production host driver is C (burst_b200/host/).  There
is no fallback: if the CUDA library is missing or no device is present, construction raises.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libburst_b200.so")

HIT_DTYPE = np.dtype([("task", "<u4"), ("lane", "u1"), ("ed", "u1"), ("gap_q", "u1"),
                      ("gap_r", "u1"), ("final_pos", "<u4")])
XHIT_DTYPE = np.dtype([("query", "<u4"), ("clump", "<u4"), ("lane", "u1"), ("ed", "u1"), ("gap_q", "u1"),
                       ("gap_r", "u1"), ("final_pos", "<u4")])
TASK_DTYPE = np.dtype([("query", "<u4"), ("clump", "<u4")])
RUN_DTYPE = np.dtype([("clump", "<u4"), ("query0", "<u4"), ("nq", "<u4")])
RUN_MAX = 16
MODE_MIN, MODE_ALL = 0, 1
Q_PACKED4 = 1
R_PACKED4, R_PACKED2 = 1, 2
PARAM_SEED_FILTER, PARAM_SEED_CHUNK, PARAM_SEED_WORDS, PARAM_SEED_STAGE, PARAM_PIPE_SLICES = 1, 2, 3, 4, 5
PARAM_PIPE_MIN_RUNS, PARAM_PIPE_RATIO, PARAM_SEED_GROUPS = 6, 7, 8
PARAM_SEED_IMPL, PARAM_SEED_NCH, PARAM_SEED_LBITS, PARAM_SEED_FB, PARAM_SEED_HSLOTS, PARAM_SEED_VMODE = 9, 10, 11, 12, 13, 14


class BgQueries(C.Structure):
    _fields_ = [("codes", C.c_void_p), ("offset", C.c_void_p), ("budget", C.c_void_p),
                ("slot", C.c_void_p), ("nq", C.c_uint32), ("nslots", C.c_uint32), ("flags", C.c_uint32)]


class BgReads(C.Structure):
    _fields_ = [("reads", C.c_void_p), ("len", C.c_void_p), ("budget", C.c_void_p), ("strand", C.c_void_p),
                ("nreads", C.c_uint32), ("nq", C.c_uint32), ("flags", C.c_uint32)]


class BgStats(C.Structure):
    _fields_ = [("tasks", C.c_uint64), ("nominal_cells", C.c_uint64), ("filter_cells", C.c_uint64),
                ("seed_steps", C.c_uint64),
                ("survivors", C.c_uint64), ("band_cells", C.c_uint64), ("hits", C.c_uint64),
                ("seed_queries", C.c_uint32), ("seed_stride", C.c_uint32), ("seed_window", C.c_uint32), ("seed_words", C.c_uint32),
                ("ms_filter", C.c_float), ("ms_extend", C.c_float), ("ms_select", C.c_float)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = ["bg_init", "bg_free", "bg_last_error", "bg_set_stream", "bg_set_scoring", "bg_default_scoring",
           "bg_load_db", "bg_batch_upload", "bg_batch_run", "bg_batch_run_extend", "bg_batch_best_device",
           "bg_batch_run_select", "bg_batch_count", "bg_batch_download", "bg_batch_stats",
           "bg_align_batch", "bg_free_hits", "bg_batch_upload_runs", "bg_align_runs", "bg_set_param", "bg_align_runs_into",
           "bg_host_alloc", "bg_host_free", "bg_align_bunches_into", "bg_stream", "bg_set_surv_cap",
           "bg_load_acx", "bg_search_bunches_into", "bg_share_db"]


def load_library(path=None):
    """Load the engine.  `path` exists for the CPU test tier only (tests/ point it at the oracle-backed
    stand-in oracle/_sim/libburst_b200_sim.so to exercise host-side logic without a GPU); the product
    always loads the CUDA library next to this file and fails loudly when it is missing."""
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the DP path)" % path)
    L = C.CDLL(path)
    L.bg_last_error.restype = C.c_char_p
    L.bg_batch_best_device.restype = C.c_void_p
    L.bg_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.bg_free.argtypes = [C.c_void_p]
    L.bg_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.bg_set_scoring.argtypes = [C.c_void_p, C.c_void_p]
    L.bg_set_param.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.bg_default_scoring.argtypes = [C.c_int, C.c_void_p]
    L.bg_load_db.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.bg_batch_upload.argtypes = [C.c_void_p, C.POINTER(BgQueries), C.c_void_p, C.c_uint64]
    L.bg_batch_run.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.bg_batch_run_extend.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.bg_batch_best_device.argtypes = [C.c_void_p]
    L.bg_batch_run_select.argtypes = [C.c_void_p, C.c_int]
    L.bg_batch_count.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.bg_batch_download.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.bg_batch_stats.argtypes = [C.c_void_p, C.POINTER(BgStats)]
    L.bg_align_batch.argtypes = [C.c_void_p, C.POINTER(BgQueries), C.c_void_p, C.c_uint64, C.c_int,
                                 C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.bg_batch_upload_runs.argtypes = [C.c_void_p, C.POINTER(BgQueries), C.c_void_p, C.c_uint64]
    L.bg_align_runs.argtypes = [C.c_void_p, C.POINTER(BgQueries), C.c_void_p, C.c_uint64, C.c_int,
                                C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.bg_free_hits.argtypes = [C.c_void_p]
    L.bg_align_runs_into.argtypes = [C.c_void_p, C.POINTER(BgQueries), C.c_void_p, C.c_uint64, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.bg_align_bunches_into.argtypes = [C.c_void_p, C.POINTER(BgReads), C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.bg_host_alloc.restype = C.c_void_p
    L.bg_stream.restype = C.c_void_p
    L.bg_stream.argtypes = [C.c_void_p]
    L.bg_set_surv_cap.argtypes = [C.c_void_p, C.c_uint32]
    L.bg_host_alloc.argtypes = [C.c_uint64]
    L.bg_load_acx.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_uint32]
    L.bg_search_bunches_into.argtypes = [C.c_void_p, C.POINTER(BgReads), C.c_uint32, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.bg_host_free.argtypes = [C.c_void_p]
    L.bg_share_db.argtypes = [C.c_void_p, C.c_void_p]
    return L


def default_scoring(z=1):
    S = np.zeros(256, np.uint8)
    load_library().bg_default_scoring(z, S.ctypes.data)
    return S


class Engine:
    def __init__(self, device=0, stream=None, lib_path=None):
        self.lib = load_library(lib_path)
        self.ctx = C.c_void_p()
        self._check(self.lib.bg_init(device, C.byref(self.ctx)))
        if stream is not None:
            self._check(self.lib.bg_set_stream(self.ctx, C.c_void_p(stream)))
        self._keep = []

    def _check(self, rc):
        if rc:
            raise RuntimeError("burst_b200: %s (code %d)" % (self.lib.bg_last_error().decode(), rc))

    def close(self):
        if self.ctx:
            self.lib.bg_free(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_scoring(self, S):
        S = np.ascontiguousarray(S, np.uint8)
        self._check(self.lib.bg_set_scoring(self.ctx, S.ctypes.data))

    def set_seed_filter(self, on):
        self._check(self.lib.bg_set_param(self.ctx, PARAM_SEED_FILTER, 1 if on else 0))

    def set_param(self, what, value):
        self._check(self.lib.bg_set_param(self.ctx, what, value))

    def load_db(self, packed, clump_len, first_clump=0):
        packed = np.ascontiguousarray(packed, np.uint8)
        clump_len = np.ascontiguousarray(clump_len, np.uint32)
        self._check(self.lib.bg_load_db(self.ctx, packed.ctypes.data, clump_len.ctypes.data,
                                        len(clump_len), first_clump))

    def share_db(self, other):
        """bg_share_db: use the database (and accelerator) `other` holds on the same device -- a second context for a second batch in flight.
        `other` must stay open for as long as this engine is."""
        self._check(self.lib.bg_share_db(self.ctx, other.ctx))
        self._keep.append(other)

    @staticmethod
    def pack4(codes):
        """BG_Q_PACKED4 form of a code array: two bases per byte, even base in the low nibble."""
        c = np.ascontiguousarray(codes, np.uint8)
        if len(c) & 1:
            c = np.concatenate([c, np.zeros(1, np.uint8)])
        return (c[0::2] | (c[1::2] << 4)).astype(np.uint8)

    @staticmethod
    def pack2(codes):
        """BG_R_PACKED2 form of a code array holding only A/C/G/T (1..4): code - 1, four bases per byte, first base in the low bits."""
        c = np.ascontiguousarray(codes, np.uint8)
        assert c.min() >= 1 and c.max() <= 4, "2-bit packing takes plain bases only"
        c = c - 1
        pad = (-len(c)) % 4
        if pad:
            c = np.concatenate([c, np.zeros(pad, np.uint8)])
        c = c.reshape(-1, 4)
        return (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)

    def align_bunches_into(self, reads, rlen, rbudget, strand, qbunch, cand_off, cand, hits_out, best_inout=None, mode=MODE_MIN, packed2=False):
        """bg_align_bunches_into: `reads` is the packed stream (pack2 / pack4 of the concatenated read codes), `strand` the sorted
        strands (read | rc << 31), (cand_off, cand) the bunch -> candidate lists.  Returns the number of hits."""
        rlen = np.ascontiguousarray(rlen, np.uint16); rbudget = np.ascontiguousarray(rbudget, np.uint16)
        strand = np.ascontiguousarray(strand, np.uint32); cand_off = np.ascontiguousarray(cand_off, np.uint32); cand = np.ascontiguousarray(cand, np.uint32)
        reads = np.ascontiguousarray(reads, np.uint8)
        R = BgReads(reads.ctypes.data, rlen.ctypes.data, rbudget.ctypes.data, strand.ctypes.data, len(rlen), len(strand), R_PACKED2 if packed2 else R_PACKED4)
        self._nslots = len(rlen)
        nh = C.c_uint64(0)
        self._check(self.lib.bg_align_bunches_into(self.ctx, C.byref(R), qbunch, cand_off.ctypes.data, cand.ctypes.data, len(cand_off) - 1, mode,
                                                   None if best_inout is None else best_inout.ctypes.data, hits_out.ctypes.data, len(hits_out), C.byref(nh)))
        return int(nh.value)

    def load_acx(self, lens, postings, word_len, big, bad):
        """bg_load_acx: the k-mer accelerator in its on-disk form (per-word posting counts, packed postings, BadList)."""
        lens = np.ascontiguousarray(lens, np.uint32); postings = np.ascontiguousarray(postings, np.uint8); bad = np.ascontiguousarray(bad, np.uint32)
        assert len(lens) == 1 << (2 * word_len)
        self._check(self.lib.bg_load_acx(self.ctx, lens.ctypes.data, postings.ctypes.data, C.c_uint64(len(postings)), word_len, int(big), bad.ctypes.data, len(bad)))

    def search_bunches_into(self, reads, rlen, rbudget, strand, qbunch, hits_out, best_inout=None, mode=MODE_MIN, heuristic=False, skip_bad=False):
        """bg_search_bunches_into: candidates from the loaded accelerator on the device, then the alignment; `reads` = pack2 of the read
        codes, `hits_out` an XHIT_DTYPE array.  Returns the number of hits."""
        rlen = np.ascontiguousarray(rlen, np.uint16); rbudget = np.ascontiguousarray(rbudget, np.uint16)
        strand = np.ascontiguousarray(strand, np.uint32); reads = np.ascontiguousarray(reads, np.uint8)
        assert hits_out.dtype == XHIT_DTYPE
        R = BgReads(reads.ctypes.data, rlen.ctypes.data, rbudget.ctypes.data, strand.ctypes.data, len(rlen), len(strand), R_PACKED2)
        self._nslots = len(rlen)
        nh = C.c_uint64(0)
        self._check(self.lib.bg_search_bunches_into(self.ctx, C.byref(R), qbunch, int(heuristic), int(skip_bad), mode,
                                                    None if best_inout is None else best_inout.ctypes.data, hits_out.ctypes.data, C.c_uint64(len(hits_out)), C.byref(nh)))
        return int(nh.value)

    def _queries(self, codes, offset, budget, slot, nslots):
        """`codes` is a uint8 code array, or a ("packed4", array) pair holding the nibble-packed form."""
        flags = 0
        if isinstance(codes, tuple):
            assert codes[0] == "packed4"
            flags, codes = Q_PACKED4, codes[1]
        codes = np.ascontiguousarray(codes, np.uint8)
        offset = np.ascontiguousarray(offset, np.uint64)
        budget = np.ascontiguousarray(budget, np.uint16)
        nq = len(offset) - 1
        if slot is None:
            slot = np.arange(nq, dtype=np.uint32)
            nslots = nq
        slot = np.ascontiguousarray(slot, np.uint32)
        q = BgQueries(codes.ctypes.data, offset.ctypes.data, budget.ctypes.data, slot.ctypes.data, nq, nslots, flags)
        self._keep = [codes, offset, budget, slot]
        return q

    @staticmethod
    def _tasks(tasks):
        if tasks is None:
            return None, 0, None
        t = np.ascontiguousarray(tasks, TASK_DTYPE) if getattr(tasks, "dtype", None) == TASK_DTYPE \
            else np.ascontiguousarray(np.asarray(tasks, np.uint32).reshape(-1, 2)).view(TASK_DTYPE).reshape(-1)
        return t, len(t), t.ctypes.data

    # ---- resident three-step form ----
    def upload(self, codes, offset, budget, tasks, slot=None, nslots=0):
        q = self._queries(codes, offset, budget, slot, nslots)
        t, n, p = self._tasks(tasks)
        self._nslots = q.nslots
        self._check(self.lib.bg_batch_upload(self.ctx, C.byref(q), p, n))

    def upload_runs(self, codes, offset, budget, runs, slot=None, nslots=0):
        q = self._queries(codes, offset, budget, slot, nslots)
        r = np.ascontiguousarray(runs, RUN_DTYPE)
        self._nslots = q.nslots
        self._check(self.lib.bg_batch_upload_runs(self.ctx, C.byref(q), r.ctypes.data, len(r)))

    def run(self, mode=MODE_MIN, best_in=None):
        b = None if best_in is None else np.ascontiguousarray(best_in, np.uint16)
        self._check(self.lib.bg_batch_run(self.ctx, mode, None if b is None else b.ctypes.data))

    def run_extend(self, mode=MODE_MIN, best_in=None):
        b = None if best_in is None else np.ascontiguousarray(best_in, np.uint16)
        self._check(self.lib.bg_batch_run_extend(self.ctx, mode, None if b is None else b.ctypes.data))

    def set_surv_cap(self, cap):
        self._check(self.lib.bg_set_surv_cap(self.ctx, cap))

    def best_device_ptr(self):
        return self.lib.bg_batch_best_device(self.ctx)

    def run_select(self, mode=MODE_MIN):
        self._check(self.lib.bg_batch_run_select(self.ctx, mode))

    def count(self):
        n = C.c_uint64(0)
        self._check(self.lib.bg_batch_count(self.ctx, C.byref(n)))
        return int(n.value)

    def download(self):
        n = self.count()
        hits = np.zeros(n, HIT_DTYPE)
        best = np.zeros(self._nslots, np.uint16)
        self._check(self.lib.bg_batch_download(self.ctx, hits.ctypes.data, n, best.ctypes.data))
        return hits, best

    def stats(self):
        s = BgStats()
        self._check(self.lib.bg_batch_stats(self.ctx, C.byref(s)))
        return s.asdict()

    def align_runs_into(self, codes, offset, budget, runs, hits_out, best_inout=None, mode=MODE_MIN, slot=None, nslots=0):
        """One call, caller-owned buffers (bg_align_runs_into): `hits_out` is a HIT_DTYPE array (pinned for full speed),
        `best_inout` an optional uint16 array of per-slot minima carried in and out.  Returns the number of hits."""
        q = self._queries(codes, offset, budget, slot, nslots)
        self._nslots = q.nslots
        r = np.ascontiguousarray(runs, RUN_DTYPE)
        nh = C.c_uint64(0)
        self._check(self.lib.bg_align_runs_into(self.ctx, C.byref(q), r.ctypes.data, len(r), mode,
                                                None if best_inout is None else best_inout.ctypes.data,
                                                hits_out.ctypes.data, len(hits_out), C.byref(nh)))
        return int(nh.value)

    # ---- one call, host buffers in, host buffers out ----
    def align(self, codes, offset, budget, tasks, mode=MODE_MIN, slot=None, nslots=0, best=None, runs=None):
        q = self._queries(codes, offset, budget, slot, nslots)
        self._nslots = q.nslots
        b = np.full(q.nslots, 0xFFFF, np.uint16) if best is None else np.ascontiguousarray(best, np.uint16).copy()
        hp = C.c_void_p(); nh = C.c_uint64(0)
        if runs is not None:
            r = np.ascontiguousarray(runs, RUN_DTYPE)
            self._check(self.lib.bg_align_runs(self.ctx, C.byref(q), r.ctypes.data, len(r), mode, b.ctypes.data, C.byref(hp), C.byref(nh)))
        else:
            t, n, p = self._tasks(tasks)
            self._check(self.lib.bg_align_batch(self.ctx, C.byref(q), p, n, mode, b.ctypes.data, C.byref(hp), C.byref(nh)))
        hits = np.zeros(nh.value, HIT_DTYPE)
        if nh.value:
            C.memmove(hits.ctypes.data, hp, nh.value * HIT_DTYPE.itemsize)
        self.lib.bg_free_hits(hp)
        return hits, b
