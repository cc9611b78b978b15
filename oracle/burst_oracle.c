/* oracle/burst_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain scalar C restatement of the arithmetic of BURST's alignment hot path, one
 * (query, reference lane) pair at a time, full matrix, no band bookkeeping, no SIMD.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (burst_b200/) never does.
 *
 * What it restates (all citations are /root/reference/burst.c):
 *   score table        setScore / SCOREFAST                         1237-1329 (rule: 172-190)
 *   char -> code       CHAR2NUM                                     1288-1307
 *   error budget       float32 (1/THRES-1)*len, capped at 254       3069-3076
 *   pass 1             aded_mat16 / aded_mat16L / ADED_PROTOTYPE    1003-1204
 *   pass 2             reScoreM_mat16 / RESCOREM_PROTYPE            713-886
 *   clump unpack       2 positions per byte per lane                4141-4150 (layout 2810-2824)
 *   task walk          Emac tightening, lanes kept at the minimum   4157-4277, 4429-4478
 *
 * Pinning: tests/test_oracle_vs_reference.py checks every function below against the
 * UNMODIFIED reference kernels (oracle/_ref/libburstref.so, built by oracle/Makefile from
 * /root/reference/burst.c) when that library is present, and tests/test_oracle_golden.py
 * checks it against the tests/golden vectors (npz), vectors produced by the same reference kernels with
 * scripts/make_golden.py.  The reference itself ships no tests or golden vectors for this
 * path.
 *
 * Semantics kept from the reference: all DP values are unsigned 8-bit saturating
 * (_mm_adds_epu8 / _mm_min_epu8); any cell whose score is >= maxED+1 is replaced by 255
 * (burst.c:1053-1054, 802-803).  The reference additionally skips cells outside a per-row
 * active range; those cells are all > maxED, so a full-matrix evaluation
 * with the same clamp yields identical values for every lane whose distance is <= maxED.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

static inline uint32_t sat8(uint32_t v) { return v > 255 ? 255 : v; }

/* ---- tables ------------------------------------------------------------------------- */

/* IUPAC code -> set of bases (bit0 A, bit1 C, bit2 G, bit3 T).  Code order follows the
 * reference's alphabet ". A C G T N K M R Y S W B V H D" (burst.c:166). */
static const uint8_t IUPAC_SET[16] = {
	0, 1, 2, 4, 8, 15, /*K=GT*/12, /*M=AC*/3, /*R=AG*/5, /*Y=CT*/10, /*S=CG*/6, /*W=AT*/9,
	/*B=CGT*/14, /*V=ACG*/7, /*H=ACT*/11, /*D=AGT*/13 };

/* S[q*16+r]: 0 when one code's base set contains the other's, else 1; every pair that
 * involves N (code 5) costs z (1 by default, 0 with -y); code 0 (pad / non-letter) costs 255
 * on either side (burst.c:1310-1328 and the 16x16 layout at 172-190). */
EXPORT void oracle_score_table(int z, uint8_t S[256]) {
	for (int q = 0; q < 16; ++q) for (int r = 0; r < 16; ++r) {
		uint8_t v;
		if (!q || !r) v = 255;
		else if (q == 5 || r == 5) v = (uint8_t)z;
		else {
			uint8_t a = IUPAC_SET[q], b = IUPAC_SET[r], i = a & b;
			v = (i == a || i == b) ? 0 : 1;
		}
		S[q * 16 + r] = v;
	}
}

/* burst.c:1288-1307: letters default to N (5), ACGTU and the ten two/three-base codes get
 * their own number, everything else (and, by the reference's loop bound, 'z') is 0. */
EXPORT void oracle_char2num(uint8_t T[128]) {
	memset(T, 0, 128);
	for (int c = 'A'; c <= 'Z'; ++c) T[c] = 5;
	for (int c = 'a'; c < 'z'; ++c) T[c] = 5;
	static const char *letters = "ACGTNKMRYSWBVHD";
	for (int i = 0; letters[i]; ++i) {
		if (letters[i] == 'N') continue;
		T[(int)letters[i]] = T[(int)letters[i] + 32] = (uint8_t)(i + 1);
	}
	T['U'] = T['u'] = 4;
}

/* burst.c:168 */
EXPORT void oracle_rc_table(uint8_t T[16]) {
	static const uint8_t rvt[16] = {0,4,3,2,1,5,7,6,9,8,10,11,13,12,15,14};
	memcpy(T, rvt, 16);
}

/* burst.c:3069-3076, all in float32 */
EXPORT uint32_t oracle_budget(float thres, uint32_t len) {
	float req = 1 / thres - 1;
	uint32_t ed = (uint32_t)(req * (float)len);
	return ed > 254 ? 254 : ed;
}

/* ---- one lane ----------------------------------------------------------------------- */

typedef struct { uint8_t sc, sh, shr; } Cell;

/* Pass 1 for one lane: semi-global unit-cost distance of q[0..m) against r[0..n) (query end
 * to end, free reference flanks).  Returns the distance if it is <= maxED, else 255. */
EXPORT uint32_t oracle_lane_ed(const uint8_t *r, uint32_t n, const uint8_t *q, uint32_t m,
		const uint8_t S[256], uint32_t maxED) {
	uint8_t *prev = calloc(n + 1, 1), *cur = malloc(n + 1);      /* row 0 is all zero */
	uint32_t bad = maxED + 1;
	for (uint32_t y = 1; y <= m; ++y) {
		cur[0] = (uint8_t)sat8(y);                               /* column 0, burst.c:1013 */
		const uint8_t *Sq = S + 16 * q[y - 1];
		for (uint32_t x = 1; x <= n; ++x) {
			uint32_t v = sat8(prev[x - 1] + Sq[r[x - 1]]);         /* burst.c:1021 */
			uint32_t u = sat8(prev[x] + 1), l = sat8(cur[x - 1] + 1);
			v = v < u ? v : u; v = v < l ? v : l;                /* 1022-1025 */
			if (v >= bad) v = 255;                               /* 1053-1054 */
			cur[x] = (uint8_t)v;
		}
		if (cur[0] >= bad) cur[0] = 255;
		uint8_t *t = prev; prev = cur; cur = t;
	}
	uint32_t best = 255;
	for (uint32_t x = 1; x <= n; ++x) if (prev[x] < best) best = prev[x];  /* 1078-1083 */
	free(prev); free(cur);
	return best <= maxED ? best : 255;
}

/* Pass 2 for one lane (burst.c:713-886): the same recurrence carrying (score, shift, shiftR)
 * with the reference's fixed tie-break order diag -> up -> left.  out = {ed, numGapQ(shift),
 * numGapR(shiftR), finalPos (1-based end column)}.  Returns ed (255 if > maxED). */
EXPORT uint32_t oracle_lane_rescore(const uint8_t *r, uint32_t n, const uint8_t *q, uint32_t m,
		const uint8_t S[256], uint32_t maxED, uint32_t out[4]) {
	Cell *prev = calloc(n + 1, sizeof(Cell)), *cur = calloc(n + 1, sizeof(Cell));
	uint32_t bad = maxED + 1 > 255 ? 255 : maxED + 1;
	for (uint32_t y = 1; y <= m; ++y) {
		cur[0].sc = (uint8_t)sat8(y); cur[0].sh = 0; cur[0].shr = (uint8_t)sat8(y);     /* 747-750 */
		const uint8_t *Sq = S + 16 * q[y - 1];
		for (uint32_t x = 1; x <= n; ++x) {
			Cell d = prev[x - 1], u = prev[x], l = cur[x - 1], c;
			if (y == 1) {   /* "Iteration 1 only", 722-739: diagonal term alone; a left shift of 1 is
			                 * recorded when the cell is a mismatch next to a zero; shiftR is 0 */
				c.sc = Sq[r[x - 1]]; c.sh = (c.sc == 1 && l.sc == 0); c.shr = 0;
				cur[x] = c; continue;
			}
			uint32_t score = sat8(d.sc + Sq[r[x - 1]]);                               /* 763-767 */
			uint32_t shift = d.sh, shiftR = d.shr;
			uint32_t scoreU = sat8(u.sc + 1), shiftU = u.sh, shiftRU = sat8(u.shr + 1); /* 768-770 */
			/* keep diag iff score <= scoreU and not (scoreU == score and shiftU > shift): 771-779 */
			if (!(score <= scoreU && !(scoreU == score && shiftU > shift)))
				shift = shiftU, shiftR = shiftRU;
			score = score < scoreU ? score : scoreU;
			uint32_t scoreL = sat8(l.sc + 1), shiftL = sat8(l.sh + 1), shiftRL = l.shr; /* 783-788 */
			if (!(score <= scoreL && !(scoreL == score && shiftL > shift)))              /* 789-798 */
				shift = shiftL, shiftR = shiftRL;
			score = score < scoreL ? score : scoreL;
			if (score >= bad) score = 255;                        /* 802-803 */
			c.sc = (uint8_t)score; c.sh = (uint8_t)shift; c.shr = (uint8_t)shiftR;
			cur[x] = c;
		}
		Cell *t = prev; prev = cur; cur = t;
	}
	/* last-row selection, left to right (826-842): replace the incumbent when the new score is
	 * smaller, or equal with a larger shift.  Incumbent starts at (255, 0, 0). */
	uint32_t bs = 255, bsh = 0, bshr = 0;
	for (uint32_t x = 1; x <= n; ++x) {
		Cell c = prev[x];
		if (c.sc < bs || (c.sc == bs && c.sh > bsh)) bsh = c.sh, bshr = c.shr;
		if (c.sc < bs) bs = c.sc;
	}
	uint32_t fp = (uint32_t)-1;                                   /* 863-879: last matching column */
	for (uint32_t x = 1; x <= n; ++x) if (prev[x].sc == bs && prev[x].sh == bsh) fp = x;
	out[0] = bs; out[1] = bsh; out[2] = bshr; out[3] = fp;
	free(prev); free(cur);
	return bs <= maxED ? bs : 255;
}

/* BLAST-style identity exactly as burst.c:844-860 computes it: one float divide, one float
 * subtract (IEEE single; compile this file without -ffast-math). */
EXPORT float oracle_identity(uint32_t ed, uint32_t qlen, uint32_t numGapQ) {
	float sc = (float)ed, den = (float)qlen + (float)numGapQ;
	return 1.0f - sc / den;
}

/* ---- one clump (16 lanes) ----------------------------------------------------------- */

/* .edx packing (burst.c:2810-2824, unpack 4141-4150): vector v holds positions 2v (low
 * nibble) and 2v+1 (high nibble); byte k of a vector belongs to lane k. */
EXPORT void oracle_unpack_clump(const uint8_t *packed, uint32_t clumplen, uint8_t *lanes /* 16*clumplen */) {
	for (uint32_t x = 0; x < clumplen; ++x)
		for (int z = 0; z < 16; ++z) {
			uint8_t b = packed[(size_t)(x >> 1) * 16 + z];
			lanes[(size_t)z * clumplen + x] = (x & 1) ? (b >> 4) : (b & 15);
		}
}

/* Both passes for one (query, clump) task.  mins[z] = pass-1 result per lane under `emac`.
 * If the minimum is <= emac, pass 2 runs with maxED = rescore_ed (0xFFFFFFFF: the minimum,
 * burst.c:4219-4227) and fills res[z*4..] = {ed, numGapQ, numGapR, finalPos} for each lane.
 * Returns the minimum (255 if no lane is within emac). */
EXPORT uint32_t oracle_task(const uint8_t *packed, uint32_t clumplen, const uint8_t *q, uint32_t m,
		const uint8_t S[256], uint32_t emac, uint32_t rescore_ed, uint8_t mins[16], uint32_t res[64]) {
	uint8_t *lanes = malloc((size_t)16 * clumplen);
	oracle_unpack_clump(packed, clumplen, lanes);
	uint32_t min = 255;
	for (int z = 0; z < 16; ++z) {
		mins[z] = (uint8_t)oracle_lane_ed(lanes + (size_t)z * clumplen, clumplen, q, m, S, emac);
		if (mins[z] < min) min = mins[z];
	}
	if (min <= emac) {
		uint32_t red = rescore_ed == 0xFFFFFFFFu ? min : rescore_ed;
		for (int z = 0; z < 16; ++z)
			oracle_lane_rescore(lanes + (size_t)z * clumplen, clumplen, q, m, S, red, res + 4 * z);
	}
	free(lanes);
	return min;
}

/* ---- a task list -------------------------------------------------------------------- */

typedef struct { uint32_t task; uint8_t lane, ed, gap_q, gap_r; uint32_t final_pos; } OracleHit;

/* The end state the reference's drivers reach for a set of (query, clump) visits
 * (burst.c:4157-4277 / 4429-4478 plus the purge rules 4219-4223, 4497-4517): every query
 * ends with the lanes, over all visited clumps, whose distance equals the per-slot minimum
 * (mode 0), or with all lanes within budget (mode 1, FORAGE).  Queries that share a slot
 * (forward / reverse-complement copies, burst.c:4218) share the running minimum.
 * Output hits are ordered by (task, lane).  Returns the number of hits (<= cap written). */
EXPORT uint64_t oracle_run_tasks(const uint8_t *packed, const uint64_t *clump_off, const uint32_t *clump_len,
		const uint8_t *qcodes, const uint64_t *qoff, const uint16_t *budget, const uint32_t *slot,
		uint32_t nslots, const uint32_t *task_query, const uint32_t *task_clump, uint64_t ntasks,
		const uint8_t S[256], int mode, uint16_t *best /* nslots, in: 0xFFFF or a bound */,
		OracleHit *hits, uint64_t cap) {
	uint8_t *mins = malloc(ntasks * 16);
	#pragma omp parallel for schedule(dynamic, 16)
	for (uint64_t t = 0; t < ntasks; ++t) {
		uint32_t qi = task_query[t], c = task_clump[t], L = clump_len[c];
		uint32_t m = (uint32_t)(qoff[qi + 1] - qoff[qi]);
		uint8_t *lanes = malloc((size_t)16 * L);
		oracle_unpack_clump(packed + clump_off[c], L, lanes);
		for (int z = 0; z < 16; ++z)
			mins[t * 16 + z] = (uint8_t)oracle_lane_ed(lanes + (size_t)z * L, L, qcodes + qoff[qi], m, S, budget[qi]);
		free(lanes);
	}
	for (uint64_t t = 0; t < ntasks; ++t) {
		uint32_t s = slot[task_query[t]];
		for (int z = 0; z < 16; ++z) if (mins[t * 16 + z] != 255 && mins[t * 16 + z] < best[s]) best[s] = mins[t * 16 + z];
	}
	uint64_t n = 0;
	for (uint64_t t = 0; t < ntasks; ++t) {
		uint32_t qi = task_query[t], c = task_clump[t], L = clump_len[c], s = slot[qi];
		uint32_t m = (uint32_t)(qoff[qi + 1] - qoff[qi]);
		uint8_t *lanes = NULL;
		for (int z = 0; z < 16; ++z) {
			uint8_t e = mins[t * 16 + z];
			if (e == 255) continue;
			if (mode == 0 && e != best[s]) continue;
			if (!lanes) { lanes = malloc((size_t)16 * L); oracle_unpack_clump(packed + clump_off[c], L, lanes); }
			uint32_t out[4];
			/* pass 2 under the bound the reference would use: the minimum (mode 0) or the budget (mode 1) */
			oracle_lane_rescore(lanes + (size_t)z * L, L, qcodes + qoff[qi], m, S, mode == 0 ? e : budget[qi], out);
			if (n < cap) {
				OracleHit h = { (uint32_t)t, (uint8_t)z, (uint8_t)out[0], (uint8_t)out[1], (uint8_t)out[2], out[3] };
				hits[n] = h;
			}
			++n;
		}
		free(lanes);
	}
	free(mins);
	return n;
}
