/* oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Function-level access to the UNMODIFIED reference kernels.  This file copies
 * nothing from the reference: it #includes the reference translation unit where
 * it lies (BURST_C_PATH, normally /root/reference/burst.c) with its main()
 * renamed, and exports thin C-ABI shims that call the reference's own
 *   setScore()            burst.c:1237-1329
 *   aded_mat16L()         burst.c:1106-1204   (pass 1, accelerated path)
 *   aded_mat16()          burst.c:1097-1098   (pass 1, fallback path)
 *   reScoreM_mat16()      burst.c:890-892     (pass 2)
 * exactly the way do_alignments() sets them up (scratch sizes burst.c:4052-4065,
 * nibble unpack burst.c:4141-4150).  The result, oracle/_ref/libburstref.so, is
 * git-ignored; it pins oracle/burst_oracle.c and generates tests/golden/.
 */
#define main burst_reference_main
#include BURST_C_PATH
#undef main

typedef struct {
	DualCoil *Matrices, *ScoresEX, *ShiftsEX, *ShiftsBX, *rclump;
	uint32_t *HiBound, *LoBound;
	void *raw[5];
	uint32_t rdim, qdim;
} ShimScratch;

static ShimScratch *shim_new(uint32_t maxLenR, uint32_t maxLenQ) {
	ShimScratch *S = calloc(1, sizeof(*S));
	uint32_t rdim = maxLenR + 2, qdim = maxLenQ + 2;        /* burst.c:3658-3659 */
	S->rdim = rdim; S->qdim = qdim;
	S->Matrices = calloc_a(64, (size_t)(cacheSz + 2) * rdim * sizeof(DualCoil), &S->raw[0]);   /* 4052 */
	S->ScoresEX = malloc_a(64, (size_t)2 * rdim * sizeof(DualCoil), &S->raw[1]);
	S->ShiftsEX = malloc_a(64, (size_t)2 * rdim * sizeof(DualCoil), &S->raw[2]);
	S->ShiftsBX = malloc_a(64, (size_t)2 * rdim * sizeof(DualCoil), &S->raw[3]);
	S->rclump   = malloc_a(64, (size_t)(2 + rdim) * sizeof(DualCoil), &S->raw[4]);
	S->HiBound = calloc(qdim + 1, sizeof(uint32_t));
	S->LoBound = calloc(qdim + 1, sizeof(uint32_t));
	for (int j = 0; j < cacheSz + 2; ++j)
		S->Matrices[(size_t)j * rdim].v = _mm_set1_epi8(MIN(j * GAP, 255));                      /* 4062 */
	*S->LoBound = -1; S->LoBound[1] = 1;                                                         /* 4063 */
	return S;
}
static void shim_free(ShimScratch *S) {
	for (int i = 0; i < 5; ++i) free(S->raw[i]);
	free(S->HiBound); free(S->LoBound); free(S);
}

/* ---- exported ---------------------------------------------------------- */
__attribute__((visibility("default")))
void refshim_set_scoring(int z, int thres_unused) { Z = (char)z; setScore(); }

__attribute__((visibility("default")))
void refshim_get_tables(uint8_t score[256], uint8_t char2num[128], uint8_t rvt[16]) {
	for (int q = 0; q < 16; ++q) {
		DualCoil d; d.v = SCOREFAST[q];
		for (int r = 0; r < 16; ++r) score[q * 16 + r] = d.u8[r];
	}
	for (int i = 0; i < 128; ++i) char2num[i] = (uint8_t)CHAR2NUM[i];
	for (int i = 0; i < 16; ++i) rvt[i] = (uint8_t)RVT[i];
}

/* float32 budget rule, burst.c:3069-3076 */
__attribute__((visibility("default")))
uint32_t refshim_budget(float thres, uint32_t len) {
	float reqID = 1/thres - 1;
	uint32_t ed = reqID * len;
	return MIN(254, ed);
}

/* Unpack a .edx clump (2 positions per byte per lane) the way burst.c:4141-4150 does */
static void shim_unpack(ShimScratch *S, const uint8_t *packed, uint32_t clumplen) {
	const DualCoil *RefSlide = (const DualCoil *)packed;
	for (uint32_t w = 0; w < clumplen; w += 2) {
		__m128i org = _mm_lddqu_si128((void*)(RefSlide++));
		__m128i ex1 = _mm_and_si128(org,_mm_set1_epi8(0xF));
		__m128i ex2 = _mm_and_si128(_mm_srli_epi16(org,4),_mm_set1_epi8(0xF));
		_mm_store_si128((void*)(S->rclump+w),ex1);
		_mm_store_si128((void*)(S->rclump+w+1),ex2);
	}
}

/* One (query, clump) task through the reference's two passes.
 * packed: .edx-packed clump (ceil(clumplen/2) * 16 bytes); query: code bytes (1..15).
 * variant 0 = aded_mat16 (fallback path), 1 = aded_mat16L (accelerated path, minlen = qlen).
 * rescore_ed: maxED handed to pass 2; 0xFFFFFFFF means "use pass-1 min" (non-FORAGE rule,
 * burst.c:4219-4227).  Outputs are written only when pass 1 returns <= emac.
 * Returns pass-1 min (0xFFFFFFFF on truncation). */
__attribute__((visibility("default")))
uint32_t refshim_task(const uint8_t *packed, uint32_t clumplen, const char *query, uint32_t qlen,
		uint32_t emac, int variant, uint32_t rescore_ed, uint8_t mins[16],
		float score[16], uint32_t finalPos[16], uint8_t numGapR[16], uint8_t numGapQ[16]) {
	ShimScratch *S = shim_new(clumplen + 1, qlen);
	shim_unpack(S, packed, clumplen);
	uint32_t rlen = clumplen + 1;
	S->HiBound[1] = rlen;                                                    /* 4151 */
	DualCoil m; m.v = _mm_set1_epi8(-1);
	uint32_t min;
	if (variant) min = aded_mat16L(S->rclump, (char*)query, rlen, qlen, S->rdim, qlen, S->Matrices, 0,
		emac, 1, S->LoBound, S->HiBound, &m);
	else min = aded_mat16(S->rclump, (char*)query, rlen, qlen, S->rdim, S->Matrices, 0,
		emac, 1, S->LoBound, S->HiBound, &m);
	if (min != (uint32_t)-1) memcpy(mins, m.u8, 16); else memset(mins, 255, 16);
	if (min <= emac) {
		MetaPack MPK __attribute__((aligned(64)));
		uint32_t red = rescore_ed == (uint32_t)-1 ? min : rescore_ed;
		reScoreM_mat16(S->rclump, (char*)query, rlen, qlen, S->rdim, S->ScoresEX, S->ShiftsEX,
			S->ShiftsBX, red, 0, &MPK);
		memcpy(score, MPK.score, sizeof(MPK.score));
		memcpy(finalPos, MPK.finalPos, sizeof(MPK.finalPos));
		memcpy(numGapR, MPK.numGapR, 16); memcpy(numGapQ, MPK.numGapQ, 16);
	}
	shim_free(S);
	return min;
}

/* CPU baseline: the reference's own kernels over a task list, OpenMP over queries.
 * Tasks must be grouped by query (task_off[q]..task_off[q+1]); each query walks its candidate
 * clumps in the given order with Emac tightening exactly like burst.c:4157-4227 (non-FORAGE).
 * packed clumps: clump c starts at packed + clump_off[c].  Returns total pass-1 calls; writes
 * best_ed[q] (0xFFFF if none) and counts pass-2 calls into *n_rescore. */
__attribute__((visibility("default")))
uint64_t refshim_run_tasks(const uint8_t *packed, const uint64_t *clump_off, const uint32_t *clump_len,
		uint32_t max_clump_len, const char *qcodes, const uint64_t *qoff, const uint16_t *budget,
		uint64_t nq, const uint32_t *task_clump, const uint64_t *task_off, int threads,
		uint16_t *best_ed, uint64_t *n_rescore, uint64_t *n_hits) {
	uint64_t calls = 0, resc = 0, hits = 0;
	uint32_t maxq = 0;
	for (uint64_t q = 0; q < nq; ++q) if (qoff[q+1]-qoff[q] > maxq) maxq = qoff[q+1]-qoff[q];
	if (threads < 1) threads = 1;
	#pragma omp parallel num_threads(threads) reduction(+:calls,resc,hits)
	{
		ShimScratch *S = shim_new(max_clump_len + 1, maxq);
		#pragma omp for schedule(dynamic,64)
		for (uint64_t q = 0; q < nq; ++q) {
			uint32_t len = qoff[q+1] - qoff[q], ed = budget[q];
			const char *query = qcodes + qoff[q];
			int found = 0;
			for (uint64_t t = task_off[q]; t < task_off[q+1]; ++t) {
				uint32_t c = task_clump[t], rlen = clump_len[c] + 1;
				shim_unpack(S, packed + clump_off[c], clump_len[c]);
				S->HiBound[1] = rlen; *S->LoBound = -1; S->LoBound[1] = 1;
				DualCoil m;
				uint32_t min = aded_mat16L(S->rclump, (char*)query, rlen, len, S->rdim, len, S->Matrices, 0,
					ed, 1, S->LoBound, S->HiBound, &m);
				++calls;
				if (min <= ed) {
					ed = min; found = 1;
					MetaPack MPK __attribute__((aligned(64)));
					reScoreM_mat16(S->rclump, (char*)query, rlen, len, S->rdim, S->ScoresEX, S->ShiftsEX,
						S->ShiftsBX, min, 0, &MPK);
					++resc;
					for (int z = 0; z < 16; ++z) hits += m.u8[z] <= min;
				}
			}
			best_ed[q] = found ? ed : 0xFFFF;
		}
		shim_free(S);
	}
	*n_rescore = resc; *n_hits = hits;
	return calls;
}

/* CPU baseline in the reference's own loop shape: the accelerated driver of do_alignments
 * (burst.c:4136-4279), restated around the reference's unmodified kernels.  Queries are the
 * sorted unique query strands (NUL-terminated code strings, qoff[j] .. ), processed in bunches
 * of `qbunch` consecutive queries (burst.c:4019-4021, 4077-4078); bunch b visits the candidate
 * clumps cand[cand_off[b] .. cand_off[b+1]) in the given order, and inside a clump every query
 * of the bunch, with
 *   - Emac = the running minimum its slot has reached so far (burst.c:4159, 4220),
 *   - prefix-row reuse: the first row to recompute is 1 + the prefix shared with the query
 *     whose rows are still valid in the cached matrix, tracked with a stack of (Emac, query)
 *     because rows computed under a budget are only reusable under budgets <= it
 *     (burst.c:4185-4211), and the instant truncation signal when the shared prefix already
 *     died (burst.c:1108),
 *   - pass 2 whenever a lane reaches the running minimum (burst.c:4217-4227).
 * The per-query k-mer-count skip (burst.c:4163-4168) is not applied: the caller expands exactly
 * the same (query, clump) pairs for the GPU arm.  ed_slot is shared and unsynchronised, as
 * ShrBins[].ed is in the reference.  Returns the number of pass-1 calls. */
/* found_slot[s] = 1 once any lane of slot s reached its running minimum (ed_slot starts at the budget, so the minimum
 * alone cannot tell "no hit" from "hit at the budget").  With sample_mod != 0 every lane the reference would push as a
 * ResultPod (burst.c:4228-4238) for a query whose slot % sample_mod == 0 is also written to rec[] (query, clump, lane,
 * ed, numGapQ, numGapR, finalPos): the caller keeps those whose ed equals the slot's final minimum (burst.c:4497-4517). */
typedef struct { uint32_t query, clump; uint8_t lane, ed, gap_q, gap_r; uint32_t final_pos; } ShimHit;
__attribute__((visibility("default")))
uint64_t refshim_run_bunches(const uint8_t *packed, const uint64_t *clump_off, const uint32_t *clump_len,
		uint32_t max_clump_len, const char *qcodes, const uint64_t *qoff, const uint32_t *qlen,
		const uint32_t *slot, uint16_t *ed_slot, uint64_t nq, uint32_t qbunch,
		const uint64_t *cand_off, const uint32_t *cand, int threads,
		uint64_t *n_rescore, uint64_t *n_hits, uint64_t *n_instant,
		uint8_t *found_slot, uint32_t sample_mod, ShimHit *rec, uint64_t rec_cap, uint64_t *n_rec) {
	uint64_t calls = 0, resc = 0, hits = 0, instant = 0, nrec = 0;
	uint32_t maxq = 0;
	for (uint64_t q = 0; q < nq; ++q) if (qlen[q] > maxq) maxq = qlen[q];
	int savedCache = cacheSz;
	cacheSz = MIN((int)maxq + 2, cacheSz);                          /* burst.c:3193 */
	/* divergence of consecutive sorted queries, burst.c:3195-3201 */
	uint16_t *Div = malloc(nq * sizeof(*Div));
	uint32_t maxDiv = 1;
	Div[0] = 1;
	for (uint64_t i = 1; i < nq; ++i) {
		const char *a = qcodes + qoff[i-1], *b = qcodes + qoff[i];
		uint32_t d = 1;
		while (*a && *a++ == *b++) ++d;
		d = MIN((uint32_t)cacheSz, d); d = MIN(qlen[i-1], d);
		Div[i] = d; if (d > maxDiv) maxDiv = d;
	}
	uint64_t nb = (nq + qbunch - 1) / qbunch;
	if (threads < 1) threads = 1;
	#pragma omp parallel num_threads(threads) reduction(+:calls,resc,hits,instant)
	{
		ShimScratch *S = shim_new(max_clump_len + 1, maxq);
		uint32_t *stE = malloc((maxq + 2 + qbunch) * sizeof(*stE));
		uint64_t *stQ = malloc((maxq + 2 + qbunch) * sizeof(*stQ));
		#pragma omp for schedule(dynamic,1)
		for (uint64_t b = 0; b < nb; ++b) {
			uint64_t z = b * qbunch, bound = MIN(z + qbunch, nq);
			uint32_t minlen = (uint32_t)-1;
			for (uint64_t j = z; j < bound; ++j) if (qlen[j] < minlen) minlen = qlen[j];
			for (uint64_t ci = cand_off[b]; ci < cand_off[b+1]; ++ci) {
				uint32_t ri = cand[ci], rlen = clump_len[ri] + 1;
				shim_unpack(S, packed + clump_off[ri], clump_len[ri]);
				S->HiBound[1] = rlen;
				uint32_t sp = 0; stQ[0] = z; stE[0] = (uint32_t)-1;
				for (uint64_t j = z; j < bound; ++j) {
					uint16_t *edp = ed_slot + slot[j];
					uint32_t Emac = *edp, len = qlen[j];
					uint32_t thisDiv = j == z ? 1 : Div[j];
					char *q = (char *)qcodes + qoff[j];
					if (Emac > stE[sp]) {                      /* rows on top were computed under a smaller budget */
						while (Emac > stE[--sp]);
						thisDiv = 1;
						if (j != z && Div[j] > 1 && sp) {
							uint64_t o = stQ[sp];
							uint32_t lim = MIN(qlen[o], len) - 1;
							const char *p = qcodes + qoff[o];
							for (uint32_t w = 0; w < lim && thisDiv < maxDiv && q[w] == p[w]; ++w) ++thisDiv;
						}
					}
					sp += Emac < stE[sp];
					stQ[sp] = j; stE[sp] = Emac;
					DualCoil m;
					uint32_t min = aded_mat16L(S->rclump, q, rlen, len, S->rdim, minlen, S->Matrices, 0,
						Emac, thisDiv, S->LoBound, S->HiBound, &m);
					++calls;
					instant += (min == (uint32_t)-1);   /* pass-1 calls that ended in truncation */
					if (min <= *edp) {
						*edp = min;
						MetaPack MPK __attribute__((aligned(64)));
						reScoreM_mat16(S->rclump, q, rlen, len, S->rdim, S->ScoresEX, S->ShiftsEX,
							S->ShiftsBX, min, 0, &MPK);
						++resc;
						for (int l = 0; l < 16; ++l) hits += m.u8[l] <= min;
						if (found_slot) found_slot[slot[j]] = 1;
						if (sample_mod && slot[j] % sample_mod == 0) for (int l = 0; l < 16; ++l) if (m.u8[l] <= min) {
							uint64_t ix;
							#pragma omp atomic capture
							ix = nrec++;
							if (ix < rec_cap) rec[ix] = (ShimHit){(uint32_t)j, ri, (uint8_t)l, m.u8[l], MPK.numGapQ[l], MPK.numGapR[l], MPK.finalPos[l]};
						}
					}
				}
			}
		}
		free(stE); free(stQ);
		shim_free(S);
	}
	free(Div);
	cacheSz = savedCache;
	*n_rescore = resc; *n_hits = hits; *n_instant = instant;
	if (n_rec) *n_rec = nrec;
	return calls;
}
