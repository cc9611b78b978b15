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
