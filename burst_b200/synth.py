"""Seeded synthetic inputs (references, clumps, reads) shared by tests and bench.py.

Shapes follow SURVEY.md 8(d): uniform ACGT references, reads cut from them with a fixed
number of edits (the model of the reference's read simulator embalmlets/LLsim.c:93-100:
exact edit count, optional reverse complement), optional IUPAC codes.
Codes are BURST's: 0 pad, 1 A, 2 C, 3 G, 4 T, 5 N, 6..15 K M R Y S W B V H D (burst.c:166).
"""
import numpy as np

RC_TABLE = np.array([0, 4, 3, 2, 1, 5, 7, 6, 9, 8, 10, 11, 13, 12, 15, 14], np.uint8)  # burst.c:168
ALPHABET = ".ACGTNKMRYSWBVHD"


def random_refs(n, length, rng, jitter=0, iupac_rate=0.0):
    """n uniform-random code strings of `length` (+- jitter) bases."""
    out = []
    for _ in range(n):
        L = length + (int(rng.integers(-jitter, jitter + 1)) if jitter else 0)
        r = rng.integers(1, 5, size=L, dtype=np.uint8)
        if iupac_rate:
            m = rng.random(L) < iupac_rate
            r[m] = rng.integers(5, 16, size=int(m.sum()), dtype=np.uint8)
        out.append(r)
    return out


def pack_clumps(refs):
    """Pack code strings 16 per clump in the .edx clump layout (burst.c:2810-2824):
    clump = ceil(ClumpLen/2) vectors of 16 bytes; byte k of vector v = lane k, low nibble =
    position 2v, high nibble = position 2v+1; short lanes are padded with code 0.
    Returns (packed uint8, clump_off uint64 [byte offsets], clump_len uint32)."""
    nclumps = (len(refs) + 15) // 16
    clump_len = np.zeros(nclumps, np.uint32)
    for i, r in enumerate(refs):
        clump_len[i // 16] = max(clump_len[i // 16], len(r))
    nvec = (clump_len.astype(np.uint64) + 1) // 2
    clump_off = np.zeros(nclumps, np.uint64)
    clump_off[1:] = np.cumsum(nvec[:-1] * 16)
    packed = np.zeros(int(nvec.sum() * 16), np.uint8)
    for c in range(nclumps):
        L = int(clump_len[c]); nv = int(nvec[c])
        m = np.zeros((16, nv * 2), np.uint8)
        for z in range(16):
            if c * 16 + z < len(refs):
                r = refs[c * 16 + z]
                m[z, :len(r)] = r
        v = (m[:, 0::2] | (m[:, 1::2] << 4)).T  # (nv, 16)
        o = int(clump_off[c])
        packed[o:o + nv * 16] = v.reshape(-1)
    return packed, clump_off, clump_len


def random_clumps(nclumps, clump_len, rng):
    """Uniform ACGT clumps of equal length directly in packed form (fast path for big DBs)."""
    nv = (clump_len + 1) // 2
    lo = rng.integers(1, 5, size=(nclumps, nv, 16), dtype=np.uint8)
    hi = rng.integers(1, 5, size=(nclumps, nv, 16), dtype=np.uint8)
    if clump_len & 1:
        hi[:, -1, :] = 0
    packed = (lo | (hi << 4)).reshape(-1)
    off = (np.arange(nclumps, dtype=np.uint64) * np.uint64(nv * 16))
    return packed, off, np.full(nclumps, clump_len, np.uint32)


def lane_codes(packed, clump_off, clump_len, clump, lane):
    """Unpack one lane of one clump to a code string (pads included)."""
    o = int(clump_off[clump]); L = int(clump_len[clump]); nv = (L + 1) // 2
    b = packed[o + lane:o + nv * 16:16]
    out = np.empty(nv * 2, np.uint8)
    out[0::2] = b & 15
    out[1::2] = b >> 4
    return out[:L]


def mutate(seq, n_edits, rng, p_sub=0.8, p_ins=0.1):
    """Apply exactly n_edits random edits (substitution / insertion / deletion)."""
    s = list(int(v) for v in seq)
    for _ in range(n_edits):
        u = rng.random()
        pos = int(rng.integers(0, len(s)))
        if u < p_sub:
            s[pos] = (s[pos] - 1 + int(rng.integers(1, 4))) % 4 + 1 if 1 <= s[pos] <= 4 else int(rng.integers(1, 5))
        elif u < p_sub + p_ins:
            s.insert(pos, int(rng.integers(1, 5)))
        elif len(s) > 1:
            del s[pos]
    return np.array(s, np.uint8)


def reads_from_clumps(packed, clump_off, clump_len, n, length, max_edits, rng, exact_edits=False,
                      rc_rate=0.0):
    """Cut n reads of `length` from random (clump, lane, offset) windows, mutate them, and
    return (list of code arrays, origin array [(clump, lane, offset, edits, rc)])."""
    reads, origin = [], np.zeros((n, 5), np.int64)
    nclumps = len(clump_len)
    for i in range(n):
        while True:
            c = int(rng.integers(0, nclumps)); z = int(rng.integers(0, 16))
            lane = lane_codes(packed, clump_off, clump_len, c, z)
            real = int(np.count_nonzero(lane))
            if real >= length:
                break
        o = int(rng.integers(0, real - length + 1))
        e = max_edits if exact_edits else int(rng.integers(0, max_edits + 1))
        r = mutate(lane[o:o + length], e, rng)
        rc = rng.random() < rc_rate
        if rc:
            r = RC_TABLE[r[::-1]]
        reads.append(r)
        origin[i] = (c, z, o, e, rc)
    return reads, origin


def concat_queries(reads):
    """-> (codes uint8 concatenated, offsets uint64 [n+1])."""
    off = np.zeros(len(reads) + 1, np.uint64)
    off[1:] = np.cumsum([len(r) for r in reads])
    codes = np.concatenate(reads).astype(np.uint8) if reads else np.zeros(0, np.uint8)
    return codes, off


def to_fasta(path, names, seqs):
    with open(path, "w") as f:
        for n, s in zip(names, seqs):
            f.write(">%s\n%s\n" % (n, "".join(ALPHABET[int(c)] for c in s)))


# ---------------------------------------------------------------------------------------------
# Vectorised generators for the bench-sized workloads (BASELINE.json configs[1] shape)
# ---------------------------------------------------------------------------------------------

def gather_windows(packed, nvec, clump, lane, offset, length):
    """Codes of `length` consecutive positions starting at `offset` of (clump, lane), for arrays
    of reads, from equal-length packed clumps (nvec vectors each)."""
    x = offset[:, None].astype(np.int64) + np.arange(length, dtype=np.int64)[None, :]
    byte = clump[:, None].astype(np.int64) * (nvec * 16) + (x >> 1) * 16 + lane[:, None].astype(np.int64)
    b = packed[byte]
    return np.where(x & 1, b >> 4, b & 15).astype(np.uint8)


def llsim_reads(packed, nclumps, clump_len, n, read_len, n_err, rng, rc=True, chunk=1 << 17):
    """n reads in the model of the reference's simulator embalmlets/LLsim.c:175-222: a window of
    read_len reference bases, exactly n_err edits at distinct positions, each a substitution
    (3/5), a deletion (1/5) or an insertion before the base (1/5); ~half reverse-complemented
    when rc.  Returns (codes, offsets[n+1], clump[n], lane[n], start[n], rcflag[n])."""
    nvec = (clump_len + 1) // 2
    all_codes, lens = [], []
    clump = rng.integers(0, nclumps, n, dtype=np.int64)
    lane = rng.integers(0, 16, n, dtype=np.int64)
    start = rng.integers(0, clump_len - read_len + 1, n, dtype=np.int64)
    rcflag = (rng.random(n) < 0.5) if rc else np.zeros(n, bool)
    for s in range(0, n, chunk):
        e = min(n, s + chunk); m = e - s
        w = gather_windows(packed, nvec, clump[s:e], lane[s:e], start[s:e], read_len)
        ins = np.zeros((m, read_len), np.uint8)          # symbol emitted before the base (0 = none)
        if n_err:
            pos = np.argsort(rng.random((m, read_len), dtype=np.float32), axis=1)[:, :n_err]
            typ = rng.integers(0, 5, (m, n_err))
            rows = np.repeat(np.arange(m), n_err); cols = pos.reshape(-1); t = typ.reshape(-1)
            sub = t < 3
            old = w[rows[sub], cols[sub]].astype(np.int64)
            w[rows[sub], cols[sub]] = ((old - 1 + 1 + t[sub]) % 4 + 1).astype(np.uint8)
            w[rows[t == 3], cols[t == 3]] = 0                # deleted
            i4 = t == 4
            ins[rows[i4], cols[i4]] = rng.integers(1, 5, int(i4.sum()), dtype=np.uint8)
        both = np.stack([ins, w], axis=2).reshape(m, -1)
        keep = both != 0
        lens.append(keep.sum(1))
        all_codes.append(both[keep])
    lens = np.concatenate(lens).astype(np.uint64)
    off = np.zeros(n + 1, np.uint64); off[1:] = np.cumsum(lens)
    codes = np.concatenate(all_codes)
    if rc:
        rcodes = reverse_complement_all(codes, off)
        sel = np.repeat(rcflag, lens.astype(np.int64))
        codes = np.where(sel, rcodes, codes)
    return codes, off, clump, lane, start, rcflag


def reverse_complement_all(codes, off):
    """Reverse-complement every read of a concatenated (codes, offsets) set (burst.c:3095-3103)."""
    lens = np.diff(off).astype(np.int64)
    rid = np.repeat(np.arange(len(lens)), lens)
    o = off[:-1].astype(np.int64)
    idx = np.arange(len(codes), dtype=np.int64)
    src = o[rid] + (lens[rid] - 1 - (idx - o[rid]))
    return RC_TABLE[codes[src]]


def sort_strands(codes, off):
    """Order of the strands under strcmp on code bytes (burst.c:3021, 3181-3184)."""
    lens = np.diff(off).astype(np.int64)
    width = int(lens.max()) + 1
    mat = np.zeros((len(lens), width), np.uint8)
    rid = np.repeat(np.arange(len(lens)), lens)
    col = np.arange(len(codes), dtype=np.int64) - off[:-1].astype(np.int64)[rid]
    mat[rid, col] = codes
    keys = mat.view("S%d" % width).reshape(-1)
    return np.argsort(keys, kind="stable")


def bunch_workload(n_reads, read_len, n_err, db_bytes, clump_len, seed, qbunch=16, budget=None):
    """BASELINE.json configs[1] shape as the reference's accelerated driver sees it: forward and
    reverse-complement strands of every read, sorted, cut into bunches of `qbunch`
    (burst.c:4019-4021); every query of a bunch visits every candidate clump of the bunch
    (burst.c:4137-4157).  Candidates = the clump each read was cut from (what the .acx lookup
    returns for error-bounded reads on a random DB: random 15-mer hits never reach the
    len-(ed+1)*N threshold, burst.c:4091-4095)."""
    rng = np.random.default_rng(seed)
    nvec = (clump_len + 1) // 2
    nclumps = max(16, int(db_bytes // (nvec * 16)))
    packed, coff, clens = random_clumps_fast(nclumps, clump_len, np.random.default_rng(1000003))
    codes, off, clump, lane, start, rcflag = llsim_reads(packed, nclumps, clump_len, n_reads, read_len, n_err, rng)
    return _bunch_form(codes, off, clump, lane, start, rcflag, n_reads, qbunch, packed, coff, clens,
                       np.full(2 * n_reads, n_err, np.uint16) if budget is None else budget, 0)


def _bunch_form(codes, off, clump, lane, start, rcflag, n_reads, qbunch, packed, coff, clens, budget, halo):
    """Strands, bunches, candidates, tasks and runs of a read set (the common tail of the bench workloads).  Candidates of a
    bunch = for every strand that matches the database, the clump it was cut from and `halo` clumps either side."""
    nclumps = len(clens)
    rcodes = reverse_complement_all(codes, off)
    lens = np.diff(off).astype(np.int64)
    # strands: 2i = as sequenced, 2i+1 = its reverse complement; the strand equal to the DB lane matches
    scodes = np.concatenate([codes, rcodes])
    soff = np.zeros(2 * n_reads + 1, np.uint64)
    soff[1:] = np.cumsum(np.concatenate([lens, lens]))
    sread = np.concatenate([np.arange(n_reads), np.arange(n_reads)])
    smatch = np.concatenate([~rcflag, rcflag])
    order = sort_strands(scodes, soff)
    # rebuild in sorted order
    slen = np.diff(soff).astype(np.int64)[order]
    qoff = np.zeros(len(order) + 1, np.uint64); qoff[1:] = np.cumsum(slen)
    src = np.repeat(soff[:-1].astype(np.int64)[order], slen) + (np.arange(int(qoff[-1]), dtype=np.int64) - np.repeat(qoff[:-1].astype(np.int64), slen))
    qcodes = scodes[src]
    slot = sread[order].astype(np.uint32)
    match = smatch[order]
    nq = len(order)
    bunch_of = np.arange(nq) // qbunch
    pb, pc = bunch_of[match], clump[slot[match]].astype(np.int64)
    if halo:
        d = np.arange(-halo, halo + 1, dtype=np.int64)
        pc = (pc[:, None] + d[None, :]).reshape(-1); pb = np.repeat(pb, len(d))
        keep = (pc >= 0) & (pc < nclumps); pb, pc = pb[keep], pc[keep]
    pairs = np.unique(np.stack([pb, pc], 1), axis=0)   # (bunch, clump), sorted
    nb = (nq + qbunch - 1) // qbunch
    cand_off = np.zeros(nb + 1, np.uint64)
    np.add.at(cand_off, pairs[:, 0] + 1, 1)
    cand_off = np.cumsum(cand_off).astype(np.uint64)
    cand = pairs[:, 1].astype(np.uint32)
    # tasks in the reference's loop order: bunch, candidate clump, query of the bunch
    qstart = pairs[:, 0] * qbunch
    qcount = np.minimum(qbunch, nq - qstart)
    tq = np.repeat(qstart, qcount) + (np.arange(int(qcount.sum())) - np.repeat(np.cumsum(qcount) - qcount, qcount))
    tc = np.repeat(cand, qcount)
    tasks = np.stack([tq, tc], 1).astype(np.uint32)
    # the same visits as run records {clump, query0, nq}: one per (bunch, candidate clump)
    runs = np.zeros(len(cand), dtype=np.dtype([("clump", "<u4"), ("query0", "<u4"), ("nq", "<u4")]))
    runs["clump"] = cand; runs["query0"] = qstart; runs["nq"] = qcount
    # the compact form of the same batch (bg_align_bunches_into): reads as sequenced, strand = read | rc << 31 in sorted order
    strand = (sread[order].astype(np.uint32) | ((order >= n_reads).astype(np.uint32) << 31)).astype(np.uint32)
    return dict(packed=packed, clump_off=coff, clump_len=clens, qcodes=qcodes, qoff=qoff, slot=slot,
                nslots=n_reads, budget=np.asarray(budget, np.uint16), tasks=tasks, runs=runs, cand_off=cand_off, cand=cand, qbunch=qbunch,
                true_clump=clump, true_lane=lane, true_start=start, n_reads=n_reads, match=match,
                rcodes=codes, rlen=lens.astype(np.uint16), strand=strand)


def amplicon_workload(n_reads, read_len, max_edits, db_bytes, ref_len, seed, budget, qbunch=16, halo=29):
    """BASELINE.json configs[2] shape: a database of `ref_len`-base marker genes in a mutation tree (phyla 25 %, families 10 %,
    genera 5 %, 3 % and species ~1.5 % apart, substitutions only), sorted so that relatives share clumps (what the reference's
    clustering is for); reads = one jittered window of a random reference with 0..max_edits substitutions; a bunch visits the
    clumps of its reads and `halo` clumps either side (~59 visits per query for halo 29: BASELINE.md section 2).  Unlike the
    shotgun shape most lanes of a visited clump lie within the budget of most queries of the bunch."""
    rng = np.random.default_rng(seed)
    nrefs = max(4096, int(db_bytes // ((ref_len + 1) // 2)) // 16 * 16)
    def children(parents, k, frac):
        ch = np.repeat(parents, k, axis=0)
        m = rng.random(ch.shape, dtype=np.float32) < frac
        sub = rng.integers(1, 4, ch.shape, dtype=np.uint8)
        return np.where(m, (ch - 1 + sub) % 4 + 1, ch).astype(np.uint8)
    nodes = rng.integers(1, 5, (1, ref_len), dtype=np.uint8)
    for frac in (0.25, 0.10, 0.05, 0.03):
        nodes = children(nodes, 4, frac)                               # 256 genera
    refs = children(nodes, (nrefs + 255) // 256, 0.015)[:nrefs]
    rows = [r.tobytes() for r in refs]
    refs = refs[np.array(sorted(range(nrefs), key=rows.__getitem__))]
    del rows
    nclumps = nrefs // 16; nv = (ref_len + 1) // 2
    m = np.zeros((nclumps, 16, nv * 2), np.uint8); m[:, :, :ref_len] = refs.reshape(nclumps, 16, ref_len)
    packed = np.ascontiguousarray((m[:, :, 0::2] | (m[:, :, 1::2] << 4)).transpose(0, 2, 1)).reshape(-1)   # clump: nv vectors of 16 lanes
    del m
    coff = np.arange(nclumps, dtype=np.uint64) * np.uint64(nv * 16); clens = np.full(nclumps, ref_len, np.uint32)
    src = rng.integers(0, nrefs, n_reads); start = 60 + rng.integers(-5, 6, n_reads)
    reads = refs[src[:, None], start[:, None] + np.arange(read_len)[None, :]]
    e = rng.integers(0, max_edits + 1, n_reads)
    for k in range(max_edits):                                           # (positions may coincide: then a read carries fewer edits)
        on = np.nonzero(e > k)[0]; pos = rng.integers(0, read_len, len(on))
        reads[on, pos] = (reads[on, pos] - 1 + rng.integers(1, 4, len(on))) % 4 + 1
    codes = reads.reshape(-1).astype(np.uint8)
    off = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
    return _bunch_form(codes, off, (src // 16).astype(np.int64), (src % 16).astype(np.int64), start + 1, np.zeros(n_reads, bool), n_reads, qbunch,
                       packed, coff, clens, np.full(2 * n_reads, budget, np.uint16), halo)


def random_clumps_fast(nclumps, clump_len, rng):
    """Uniform ACGT clumps of one length in packed .edx clump form, from raw random bytes (in 1 GB pieces: the 31.5 GB
    target database must not need several temporaries of its own size)."""
    nv = (clump_len + 1) // 2
    total = nclumps * nv * 16
    packed = np.empty(total, np.uint8)
    step = 1 << 30
    for a in range(0, total, step):
        b = min(total, a + step)
        raw = np.frombuffer(rng.bytes(b - a), np.uint8)
        np.bitwise_and(raw, 3, out=packed[a:b]); packed[a:b] += 1
        hi = (raw >> 2) & 3; hi += 1; hi <<= 4
        packed[a:b] |= hi
    if clump_len & 1:
        packed.reshape(nclumps, nv, 16)[:, -1, :] &= 15
    off = np.arange(nclumps, dtype=np.uint64) * np.uint64(nv * 16)
    return packed, off, np.full(nclumps, clump_len, np.uint32)


# ---------------------------------------------------------------------------------------------
# Amplicon-shaped inputs (BASELINE.json configs[2]): references in a mutation tree, so that the lanes of a
# clump are near-identical and most of them tie; reads cut from a (jittered) fixed start of their reference.
# ---------------------------------------------------------------------------------------------

def mutation_tree_refs(rng, n_refs, length, levels=(0.25, 0.10, 0.05, 0.03), fanout=(4, 4, 4)):
    """n_refs references of ~`length` bases: a root sequence, `fanout[0]` phyla at levels[0] divergence from it,
    families under them at levels[1], genera at levels[2], and leaves at levels[3] from their genus
    (25 % / 10 % / 5 % / 3 %).  Substitutions only above the leaves, any edit at the leaves."""
    def diverge(seq, frac, subs_only):
        s = seq.copy()
        n = int(round(frac * len(s)))
        if subs_only:
            pos = rng.choice(len(s), n, replace=False)
            s[pos] = (s[pos] - 1 + rng.integers(1, 4, n)) % 4 + 1
            return s.astype(np.uint8)
        return mutate(s, n, rng, p_sub=0.8, p_ins=0.1)
    root = rng.integers(1, 5, length, dtype=np.uint8)
    nodes = [root]
    for lv, fo in zip(levels[:-1], fanout):
        nodes = [diverge(p, lv, True) for p in nodes for _ in range(fo)]
    out = []
    per = (n_refs + len(nodes) - 1) // len(nodes)
    for g in nodes:
        for _ in range(per):
            if len(out) < n_refs:
                out.append(diverge(g, levels[-1] * rng.random(), False))
    return out


def amplicon_reads(refs, n, read_len, max_edits, rng, start=40, jitter=5, dup_rate=0.0):
    """n reads of read_len bases starting within `jitter` of `start` on a random reference, 0..max_edits edits;
    a fraction dup_rate repeats an earlier read exactly (amplicon data is full of duplicates)."""
    reads, src = [], np.zeros(n, np.int64)
    for i in range(n):
        if reads and rng.random() < dup_rate:
            j = int(rng.integers(0, len(reads)))
            reads.append(reads[j].copy()); src[i] = src[j]
            continue
        while True:
            r = int(rng.integers(0, len(refs)))
            if len(refs[r]) >= start + jitter + read_len:
                break
        o = start + int(rng.integers(-jitter, jitter + 1))
        reads.append(mutate(refs[r][o:o + read_len], int(rng.integers(0, max_edits + 1)), rng))
        src[i] = r
    return reads, src


def bunch_runs(order_len, qbunch, cands_of_bunch):
    """Run records {clump, query0, nq} and the equivalent (query, clump) task list + hit keys for bunches of `qbunch`
    consecutive queries; cands_of_bunch(b, q0, n) -> iterable of clump ids in visiting order."""
    runs, tq, tc, key = [], [], [], []
    for b, q0 in enumerate(range(0, order_len, qbunch)):
        n = min(qbunch, order_len - q0)
        for c in cands_of_bunch(b, q0, n):
            for i in range(n):
                tq.append(q0 + i); tc.append(int(c)); key.append(len(runs) * 16 + i)
            runs.append((int(c), q0, n))
    runs = np.array(runs, dtype=np.dtype([("clump", "<u4"), ("query0", "<u4"), ("nq", "<u4")]))
    return runs, np.array(tq, np.uint32), np.array(tc, np.uint32), np.array(key, np.uint32)


def strand_batch(reads, budgets, qbunch, cands_of_bunch, rng=None):
    """The compact strand form (bg_align_bunches_into) of a read set and the equivalent general form.
    reads: list of code arrays; budgets: per read.  Strands = every read and its reverse complement, sorted as the
    reference sorts them (burst.c:3181-3184).  cands_of_bunch(b, strand_reads, strand_rc) -> candidate clumps of bunch b.
    Returns dict(rlen, rbudget, strand, cand_off, cand, rcodes [concatenated read codes], and the general form:
    qcodes, qoff, budget, slot, runs, tq, tc, key)."""
    n = len(reads)
    strands, sread, src = [], [], []
    for i, r in enumerate(reads):
        strands += [r, RC_TABLE[r[::-1]]]; sread += [i, i]; src += [0, 1]
    codes, off = concat_queries(strands)
    order = sort_strands(codes, off)
    strands = [strands[i] for i in order]
    sread = np.array([sread[i] for i in order], np.uint32); src = np.array([src[i] for i in order], np.uint32)
    nq = len(strands)
    cand_off = [0]; cand = []
    for b, q0 in enumerate(range(0, nq, qbunch)):
        cs = list(cands_of_bunch(b, sread[q0:q0 + qbunch], src[q0:q0 + qbunch]))
        cand += [int(c) for c in cs]; cand_off.append(len(cand))
    cand_off = np.array(cand_off, np.uint32); cand = np.array(cand, np.uint32)
    it = iter(range(len(cand_off) - 1))
    runs, tq, tc, key = bunch_runs(nq, qbunch, lambda b, q0, m: cand[cand_off[b]:cand_off[b + 1]])
    qcodes, qoff = concat_queries(strands)
    rcodes, _ = concat_queries(reads)
    return dict(rlen=np.array([len(r) for r in reads], np.uint16), rbudget=np.asarray(budgets, np.uint16), strand=(sread | (src << 31)).astype(np.uint32),
                cand_off=cand_off, cand=cand, rcodes=rcodes, qcodes=qcodes, qoff=qoff, budget=np.asarray(budgets, np.uint16)[sread], slot=sread,
                runs=runs, tq=tq, tc=tc, key=key, nreads=n)


def build_acx(refs, N, bad=(), big=False):
    """A k-mer accelerator over `refs` (16 per clump) in the .acx form (burst.c:3504-3530): for every N-mer of plain bases the ascending
    list of clumps holding it; clumps in `bad` are left out of the lists and named in the BadList instead.  Returns (lens uint32 [4^N],
    postings uint8 [small format: two 20-bit ids in 5 bytes, an odd last one in 3; big: 3 bytes each], bad uint32, lists dict word -> ids)."""
    bad = sorted(int(b) for b in bad)
    pairs = []
    for c in range((len(refs) + 15) // 16):
        if c in bad:
            continue
        ws = []
        for r in refs[c * 16:c * 16 + 16]:
            r = np.asarray(r, np.int64)
            if len(r) < N:
                continue
            ok = (r >= 1) & (r <= 4)
            w = np.zeros(len(r) - N + 1, np.int64); good = np.ones(len(r) - N + 1, bool)
            for t in range(N):
                w = (w << 2) | ((r[t:len(r) - N + 1 + t] - 1) & 3); good &= ok[t:len(r) - N + 1 + t]
            ws.append(w[good])
        if ws:
            u = np.unique(np.concatenate(ws))
            pairs.append(np.stack([u, np.full(len(u), c, np.int64)], 1))
    pairs = np.concatenate(pairs) if pairs else np.zeros((0, 2), np.int64)
    pairs = pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]
    lens = np.bincount(pairs[:, 0], minlength=1 << (2 * N)).astype(np.uint32)
    words, starts = np.unique(pairs[:, 0], return_index=True)
    ends = list(starts[1:]) + [len(pairs)]
    out = bytearray(); lists = {}
    for w, a, b in zip(words, starts, ends):
        ids = pairs[a:b, 1]
        lists[int(w)] = ids
        if big:
            for i in ids:
                out += int(i).to_bytes(3, "little")
        else:
            for k in range(0, len(ids) - 1, 2):
                out += (int(ids[k]) | (int(ids[k + 1]) << 20)).to_bytes(5, "little")
            if len(ids) & 1:
                out += int(ids[-1]).to_bytes(3, "little")
    return lens, np.frombuffer(bytes(out), np.uint8), np.array(bad, np.uint32), lists


def acx_runs(strands, budgets, qbunch, N, lists, bad, nclumps, heuristic=False, skip_bad=False):
    """The reference's candidate rule for bunches of `qbunch` sorted strands (burst.c:4085-4168), restated: every distinct word of the
    bunch adds its largest per-query multiplicity to each clump of its list; clumps whose count exceeds the bunch threshold are visited
    by descending count (equal counts: first touched first), each by the maximal ranges of queries whose own threshold it exceeds; then
    the BadList.  Returns the runs as a list of (clump, query0, nq)."""
    runs = []
    for z in range(0, len(strands), qbunch):
        nb = min(qbunch, len(strands) - z)
        mm, minmm, mult = [], 1 << 62, {}
        for j in range(nb):
            s = np.asarray(strands[z + j], np.int64); ln = len(s); kload = int(budgets[z + j]) * N + N
            m = ln - kload if kload < ln else 0
            if heuristic:
                m = max(m, (ln >> 4) + 1)
            minmm = min(minmm, m)
            mm.append(ln - kload if kload < ln else 1)
            per = {}
            for k in range(0, ln - N + 1):
                w = 0
                for t in range(N):
                    w = (w << 2) | (int(s[k + t]) - 1)
                per[w] = per.get(w, 0) + 1
            for w, n in per.items():
                mult[w] = max(mult.get(w, 0), n)
        count, order = {}, []
        for w in sorted(mult):
            for c in lists.get(w, ()):
                c = int(c)
                if c >= nclumps:
                    continue
                if c not in count:
                    count[c] = 0; order.append(c)
                count[c] += mult[w]
        cands = [(min(count[c], 65535), i, c) for i, c in enumerate(order) if min(count[c], 65535) > minmm]
        cands.sort(key=lambda t: (-t[0], t[1]))
        for v, _, c in cands:
            a = 0
            while a < nb:
                while a < nb and not v > mm[a]:
                    a += 1
                b = a
                while b < nb and v > mm[b]:
                    b += 1
                if b > a:
                    runs.append((c, z + a, b - a))
                a = b
        if not skip_bad:
            for c in bad:
                if int(c) < nclumps:
                    runs.append((int(c), z, nb))
    return runs
