"""ctypes bindings for the CHECKERS (test infrastructure only).

  Oracle      -> oracle/liboracle.so        the scalar C restatement (oracle/burst_oracle.c)
  Reference   -> oracle/_ref/libburstref.so the UNMODIFIED reference kernels behind
                                            oracle/ref_shim.c (present only where
                                            /root/reference was available at build time)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package burst_b200/ never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HIT_DTYPE = np.dtype([("task", "<u4"), ("lane", "u1"), ("ed", "u1"), ("gap_q", "u1"),
                      ("gap_r", "u1"), ("final_pos", "<u4")])


def build(ref=True):
    """Compile the oracle (and, when /root/reference exists, the reference) with oracle/Makefile."""
    subprocess.run(["make", "-s", "-C", HERE, "liboracle.so"], check=True)
    if ref and os.path.exists("/root/reference/burst.c"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        L.oracle_budget.restype = C.c_uint32
        L.oracle_budget.argtypes = [C.c_float, C.c_uint32]
        L.oracle_identity.restype = C.c_float
        L.oracle_identity.argtypes = [C.c_uint32] * 3
        L.oracle_task.restype = C.c_uint32
        L.oracle_lane_ed.restype = C.c_uint32
        L.oracle_lane_rescore.restype = C.c_uint32
        L.oracle_run_tasks.restype = C.c_uint64

    def score_table(self, z=1):
        S = np.zeros(256, np.uint8)
        self.lib.oracle_score_table(C.c_int(z), _p(S))
        return S

    def char2num(self):
        T = np.zeros(128, np.uint8)
        self.lib.oracle_char2num(_p(T))
        return T

    def rc_table(self):
        T = np.zeros(16, np.uint8)
        self.lib.oracle_rc_table(_p(T))
        return T

    def budget(self, thres, length):
        return int(self.lib.oracle_budget(C.c_float(thres), C.c_uint32(length)))

    def identity(self, ed, qlen, gap_q):
        return np.float32(self.lib.oracle_identity(ed, qlen, gap_q))

    def task(self, packed, clumplen, q, S, emac, rescore_ed=0xFFFFFFFF):
        packed = np.ascontiguousarray(packed, np.uint8)
        q = np.ascontiguousarray(q, np.uint8)
        mins = np.zeros(16, np.uint8)
        res = np.zeros(64, np.uint32)
        m = self.lib.oracle_task(_p(packed), C.c_uint32(clumplen), _p(q), C.c_uint32(len(q)), _p(S),
                                 C.c_uint32(emac), C.c_uint32(rescore_ed), _p(mins), _p(res))
        return int(m), mins, res.reshape(16, 4)

    def run_tasks(self, packed, clump_off, clump_len, qcodes, qoff, budget, slot, nslots,
                  task_query, task_clump, S, mode=0, best=None):
        """Returns (hits sorted by (task, lane), best[nslots])."""
        ntasks = len(task_query)
        best = np.full(nslots, 0xFFFF, np.uint16) if best is None else best.astype(np.uint16).copy()
        cap = max(1024, ntasks * 4)
        while True:
            hits = np.zeros(cap, HIT_DTYPE)
            b = best.copy()
            n = self.lib.oracle_run_tasks(
                _p(np.ascontiguousarray(packed, np.uint8)), _p(np.ascontiguousarray(clump_off, np.uint64)),
                _p(np.ascontiguousarray(clump_len, np.uint32)), _p(np.ascontiguousarray(qcodes, np.uint8)),
                _p(np.ascontiguousarray(qoff, np.uint64)), _p(np.ascontiguousarray(budget, np.uint16)),
                _p(np.ascontiguousarray(slot, np.uint32)), C.c_uint32(nslots),
                _p(np.ascontiguousarray(task_query, np.uint32)), _p(np.ascontiguousarray(task_clump, np.uint32)),
                C.c_uint64(ntasks), _p(S), C.c_int(mode), _p(b), _p(hits), C.c_uint64(cap))
            if n <= cap:
                return hits[:n], b
            cap = int(n)


class Reference:
    """The reference's own aded_mat16 / aded_mat16L / reScoreM_mat16 (burst.c) through ref_shim.c."""

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "libburstref.so"))

    def __init__(self):
        self.lib = L = C.CDLL(os.path.join(HERE, "_ref", "libburstref.so"))
        L.refshim_task.restype = C.c_uint32
        L.refshim_budget.restype = C.c_uint32
        L.refshim_budget.argtypes = [C.c_float, C.c_uint32]
        L.refshim_run_tasks.restype = C.c_uint64
        self.set_scoring(1)

    def set_scoring(self, z):
        self.lib.refshim_set_scoring(C.c_int(z), C.c_int(0))

    def tables(self):
        S = np.zeros(256, np.uint8); c2n = np.zeros(128, np.uint8); rvt = np.zeros(16, np.uint8)
        self.lib.refshim_get_tables(_p(S), _p(c2n), _p(rvt))
        return S, c2n, rvt

    def budget(self, thres, length):
        return int(self.lib.refshim_budget(C.c_float(thres), C.c_uint32(length)))

    def task(self, packed, clumplen, q, emac, variant=1, rescore_ed=0xFFFFFFFF):
        packed = np.ascontiguousarray(packed, np.uint8)
        q = np.ascontiguousarray(q, np.uint8)
        mins = np.zeros(16, np.uint8); score = np.zeros(16, np.float32)
        fp = np.zeros(16, np.uint32); gr = np.zeros(16, np.uint8); gq = np.zeros(16, np.uint8)
        m = self.lib.refshim_task(_p(packed), C.c_uint32(clumplen), _p(q), C.c_uint32(len(q)),
                                  C.c_uint32(emac), C.c_int(variant), C.c_uint32(rescore_ed),
                                  _p(mins), _p(score), _p(fp), _p(gr), _p(gq))
        return int(m), mins, score, fp, gr, gq

    def run_tasks(self, packed, clump_off, clump_len, qcodes, qoff, budget, task_clump, task_off, threads):
        nq = len(qoff) - 1
        best = np.zeros(nq, np.uint16)
        nres = C.c_uint64(0); nhits = C.c_uint64(0)
        calls = self.lib.refshim_run_tasks(
            _p(packed), _p(np.ascontiguousarray(clump_off, np.uint64)), _p(np.ascontiguousarray(clump_len, np.uint32)),
            C.c_uint32(int(np.max(clump_len))), _p(qcodes), _p(np.ascontiguousarray(qoff, np.uint64)),
            _p(np.ascontiguousarray(budget, np.uint16)), C.c_uint64(nq),
            _p(np.ascontiguousarray(task_clump, np.uint32)), _p(np.ascontiguousarray(task_off, np.uint64)),
            C.c_int(threads), _p(best), C.byref(nres), C.byref(nhits))
        return int(calls), best, int(nres.value), int(nhits.value)


def nul_terminated(qcodes, qoff):
    """Concatenated code strings -> the reference's NUL-terminated form; returns (bytes, offsets, lens)."""
    lens = np.diff(qoff).astype(np.int64)
    n = len(lens)
    out = np.zeros(int(lens.sum()) + n, np.uint8)
    noff = (qoff[:-1].astype(np.int64) + np.arange(n)).astype(np.uint64)
    rid = np.repeat(np.arange(n), lens)
    dst = np.arange(len(qcodes), dtype=np.int64) + rid
    out[dst] = qcodes
    return out, noff, lens.astype(np.uint32)


SHIMHIT_DTYPE = np.dtype([("query", "<u4"), ("clump", "<u4"), ("lane", "u1"), ("ed", "u1"), ("gap_q", "u1"), ("gap_r", "u1"), ("final_pos", "<u4")])


def reference_run_bunches(ref, w, nbunches=None, threads=1, sample_mod=0):
    """Drive Reference (libburstref.so) over the first nbunches bunches of a synth.bunch_workload.
    With sample_mod, also returns under "hits" the lanes the reference keeps (ed == final minimum of the slot) for the
    queries whose slot % sample_mod == 0, sorted by (query, clump, lane)."""
    qb = w["qbunch"]
    nq_all = len(w["qoff"]) - 1
    nb_all = (nq_all + qb - 1) // qb
    nb = nb_all if nbunches is None else min(nbunches, nb_all)
    nq = min(nq_all, nb * qb)
    qc, noff, lens = nul_terminated(w["qcodes"][:int(w["qoff"][nq])], w["qoff"][:nq + 1])
    ed = np.full(w["nslots"], 0, np.uint16)
    ed[w["slot"][:nq]] = w["budget"][:nq]          # ShrBins[].ed starts at the budget (burst.c:3076)
    budget0 = ed.copy()
    nres = C.c_uint64(0); nhits = C.c_uint64(0); ninst = C.c_uint64(0); nrec = C.c_uint64(0)
    found = np.zeros(w["nslots"], np.uint8)
    rec = np.zeros((nq // max(sample_mod, 1)) * 8 + 1024 if sample_mod else 1, SHIMHIT_DTYPE)
    L = ref.lib
    L.refshim_run_bunches.restype = C.c_uint64
    calls = L.refshim_run_bunches(
        _p(w["packed"]), _p(np.ascontiguousarray(w["clump_off"], np.uint64)),
        _p(np.ascontiguousarray(w["clump_len"], np.uint32)), C.c_uint32(int(w["clump_len"].max())),
        _p(qc), _p(noff), _p(lens), _p(np.ascontiguousarray(w["slot"][:nq], np.uint32)), _p(ed),
        C.c_uint64(nq), C.c_uint32(qb), _p(np.ascontiguousarray(w["cand_off"][:nb + 1], np.uint64)),
        _p(np.ascontiguousarray(w["cand"], np.uint32)), C.c_int(threads),
        C.byref(nres), C.byref(nhits), C.byref(ninst),
        _p(found), C.c_uint32(sample_mod), _p(rec), C.c_uint64(len(rec)), C.byref(nrec))
    assert nrec.value <= len(rec), "sample record buffer too small"
    best = np.where(found != 0, ed, 0xFFFF).astype(np.uint16)         # the ABI's convention: 0xFFFF = no lane within budget
    rec = rec[:nrec.value]
    rec = rec[rec["ed"] == best[w["slot"][rec["query"]]]]
    rec = rec[np.lexsort((rec["lane"], rec["clump"], rec["query"]))]
    return dict(calls=int(calls), rescore=int(nres.value), lanes=int(nhits.value), truncated=int(ninst.value),
                ed=ed, best=best, found=found, hits=rec, budget0=budget0, nq=nq, nb=nb)
