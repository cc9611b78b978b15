/* oracle/abi_sim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * The C ABI of include/burst_b200.h implemented over the scalar oracle (burst_oracle.c), so that
 * the HOST logic of the drop-in binary (burst_b200/host/burst_b200.c: CLI, FASTA/.edx/.acx
 * readers, query preprocessing, candidate generation, pod lists, reporters) can be exercised by
 * the CPU-only test tier (`-m "not gpu"`) in a container without a GPU.  oracle/Makefile links it
 * into oracle/_sim/burst-b200-sim.  It is never built into, linked by, or loaded from the product
 * (burst_b200/libburst_b200.so has no CPU path and bg_init fails without a CUDA device).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "burst_b200.h"

typedef struct { uint32_t task; uint8_t lane, ed, gap_q, gap_r; uint32_t final_pos; } OracleHit;
void oracle_score_table(int z, uint8_t S[256]);
uint64_t oracle_run_tasks(const uint8_t *packed, const uint64_t *clump_off, const uint32_t *clump_len,
		const uint8_t *qcodes, const uint64_t *qoff, const uint16_t *budget, const uint32_t *slot,
		uint32_t nslots, const uint32_t *task_query, const uint32_t *task_clump, uint64_t ntasks,
		const uint8_t S[256], int mode, uint16_t *best, OracleHit *hits, uint64_t cap);

struct bg_ctx {
	uint8_t S[256];
	uint8_t *packed; uint64_t *clump_off; uint32_t *clump_len; uint32_t num_clumps, first_clump; int borrowed;   /* borrowed: database + accelerator belong to another context (bg_share_db) */
	/* batch */
	uint8_t *codes; uint64_t *qoff; uint16_t *budget; uint32_t *slot; uint32_t nq, nslots;
	uint32_t *tq, *tc; uint64_t *orig; uint64_t ntasks;
	uint16_t *best; OracleHit *hits; uint64_t nhits;
	uint32_t *best32;           /* what bg_batch_best_device() exposes between run_extend and run_select */
	uint64_t *acx_off; uint8_t *acx_post; uint32_t *acx_bad, acx_nbad; int acx_n, acx_big;
};

static char g_err[256] = "";
const char *bg_last_error(void) { return g_err; }
void bg_default_scoring(int z, uint8_t S[256]) { oracle_score_table(z, S); }

int bg_init(int device, bg_ctx **out) {
	(void)device;
	bg_ctx *c = calloc(1, sizeof(*c));
	oracle_score_table(1, c->S);
	*out = c;
	return BG_OK;
}
static void free_batch(bg_ctx *c) {
	free(c->codes); free(c->qoff); free(c->budget); free(c->slot); free(c->tq); free(c->tc); free(c->orig);
	free(c->best); free(c->hits); free(c->best32); c->best32 = NULL;
	c->codes = NULL; c->qoff = NULL; c->budget = NULL; c->slot = NULL; c->tq = c->tc = NULL; c->orig = NULL; c->best = NULL; c->hits = NULL;
}
static void drop_db(bg_ctx *c) {
	if (!c->borrowed) { free(c->packed); free(c->clump_off); free(c->clump_len); free(c->acx_off); free(c->acx_post); free(c->acx_bad); }
	c->packed = NULL; c->clump_off = NULL; c->clump_len = NULL; c->acx_off = NULL; c->acx_post = NULL; c->acx_bad = NULL; c->borrowed = 0; c->num_clumps = 0; c->acx_n = 0;
}
void bg_free(bg_ctx *c) { if (!c) return; free_batch(c); drop_db(c); free(c); }
void *bg_host_alloc(uint64_t bytes) { return malloc(bytes ? bytes : 1); }
void bg_host_free(void *p) { free(p); }
int bg_set_stream(bg_ctx *c, void *s) { (void)c; (void)s; return BG_OK; }
int bg_set_param(bg_ctx *c, int what, int value) { (void)c; (void)what; (void)value; return BG_OK; }
int bg_set_scoring(bg_ctx *c, const uint8_t S[256]) { memcpy(c->S, S, 256); return BG_OK; }

int bg_load_db(bg_ctx *c, const uint8_t *packed, const uint32_t *clump_len, uint32_t n, uint32_t first) {
	if (c->borrowed) drop_db(c);
	free(c->packed); free(c->clump_off); free(c->clump_len);
	c->clump_off = malloc((n + 1) * 8); c->clump_len = malloc(n * 4);
	uint64_t tot = 0;
	for (uint32_t i = 0; i < n; ++i) { c->clump_off[i] = tot; tot += (uint64_t)((clump_len[i] + 1) / 2) * 16; c->clump_len[i] = clump_len[i]; }
	c->clump_off[n] = tot;
	c->packed = malloc(tot ? tot : 1); memcpy(c->packed, packed, tot);
	c->num_clumps = n; c->first_clump = first;
	return BG_OK;
}

static void *dup(const void *p, size_t n) { void *r = malloc(n ? n : 1); memcpy(r, p, n); return r; }

static void set_queries(bg_ctx *c, const bg_queries *Q) {
	free_batch(c);
	c->nq = Q->nq; c->nslots = Q->nslots;
	if (Q->flags & BG_Q_PACKED4) {                     /* nibble stream -> one code per byte */
		uint64_t nb = Q->offset[Q->nq];
		uint8_t *u = malloc(nb + 1);
		for (uint64_t i = 0; i < nb; ++i) u[i] = (Q->codes[i >> 1] >> (4 * (i & 1))) & 15;
		c->codes = u;
	} else c->codes = dup(Q->codes, Q->offset[Q->nq]);
	c->qoff = dup(Q->offset, (Q->nq + 1) * 8);
	c->budget = dup(Q->budget, Q->nq * 2); c->slot = dup(Q->slot, Q->nq * 4);
}

/* keeps only tasks whose clump lies in this shard; orig[] maps back to the caller's task index */
static void push_task(bg_ctx *c, uint64_t *n, uint32_t q, uint32_t clump, uint64_t orig) {
	uint32_t cl = clump - c->first_clump;
	if (cl >= c->num_clumps) return;
	c->tq[*n] = q; c->tc[*n] = cl; c->orig[*n] = orig; ++*n;
}

int bg_batch_upload(bg_ctx *c, const bg_queries *Q, const bg_task *tasks, uint64_t ntasks) {
	set_queries(c, Q);
	if (!tasks) ntasks = (uint64_t)Q->nq * c->num_clumps;
	c->tq = malloc((ntasks + 1) * 4); c->tc = malloc((ntasks + 1) * 4); c->orig = malloc((ntasks + 1) * 8);
	uint64_t n = 0;
	for (uint64_t t = 0; t < ntasks; ++t) {
		if (tasks) push_task(c, &n, tasks[t].query, tasks[t].clump, t);
		else push_task(c, &n, (uint32_t)(t % Q->nq), (uint32_t)(t / Q->nq) + c->first_clump, t);
	}
	c->ntasks = n;
	return BG_OK;
}

int bg_batch_upload_runs(bg_ctx *c, const bg_queries *Q, const bg_run *runs, uint64_t nruns) {
	set_queries(c, Q);
	uint64_t cap = nruns * BG_RUN_MAX;
	c->tq = malloc((cap + 1) * 4); c->tc = malloc((cap + 1) * 4); c->orig = malloc((cap + 1) * 8);
	uint64_t n = 0;
	for (uint64_t r = 0; r < nruns; ++r) {
		if (!runs[r].nq || runs[r].nq > BG_RUN_MAX || (uint64_t)runs[r].query0 + runs[r].nq > Q->nq) {
			snprintf(g_err, sizeof(g_err), "bg_batch_upload_runs: run %llu is malformed", (unsigned long long)r); return BG_EINVAL;
		}
		for (uint32_t i = 0; i < runs[r].nq; ++i) push_task(c, &n, runs[r].query0 + i, runs[r].clump, r * BG_RUN_MAX + i);
	}
	c->ntasks = n;
	return BG_OK;
}

static int run(bg_ctx *c, int mode, const uint16_t *best_in) {
	free(c->best); free(c->hits);
	c->best = malloc(c->nslots * 2 + 2);
	for (uint32_t i = 0; i < c->nslots; ++i) c->best[i] = best_in ? best_in[i] : 0xFFFF;
	uint64_t cap = c->ntasks * 16 + 1;
	c->hits = malloc(cap * sizeof(OracleHit));
	c->nhits = oracle_run_tasks(c->packed, c->clump_off, c->clump_len, c->codes, c->qoff, c->budget, c->slot, c->nslots,
		c->tq, c->tc, c->ntasks, c->S, mode, c->best, c->hits, cap);
	return BG_OK;
}
int bg_batch_run(bg_ctx *c, int mode, const uint16_t *best_in) { return run(c, mode, best_in); }
/* split form: extend leaves the local minima in best32 (host memory here); the caller may lower them
 * (all-reduce MIN across reference shards); select re-derives the hits under the lowered minima */
int bg_batch_run_extend(bg_ctx *c, int mode, const uint16_t *best_in) {
	run(c, mode, best_in);
	free(c->best32); c->best32 = malloc(c->nslots * 4 + 4);
	for (uint32_t i = 0; i < c->nslots; ++i) c->best32[i] = c->best[i];
	return BG_OK;
}
void *bg_batch_best_device(bg_ctx *c) { return c->best32; }
void *bg_stream(bg_ctx *c) { (void)c; return 0; }
int bg_set_surv_cap(bg_ctx *c, uint32_t cap) { (void)c; (void)cap; return BG_OK; }
int bg_batch_run_select(bg_ctx *c, int mode) {
	uint16_t *in = malloc(c->nslots * 2 + 2);
	for (uint32_t i = 0; i < c->nslots; ++i) in[i] = (uint16_t)(c->best32[i] > 0xFFFF ? 0xFFFF : c->best32[i]);
	run(c, mode, in);
	free(in);
	return BG_OK;
}
int bg_batch_count(bg_ctx *c, uint64_t *n) { if (n) *n = c->nhits; return BG_OK; }
int bg_batch_download(bg_ctx *c, bg_hit *hits, uint64_t cap, uint16_t *best_out) {
	if (hits) {
		if (cap < c->nhits) return BG_EINVAL;
		for (uint64_t i = 0; i < c->nhits; ++i) {
			OracleHit h = c->hits[i];
			bg_hit o = {(uint32_t)c->orig[h.task], h.lane, h.ed, h.gap_q, h.gap_r, h.final_pos};
			hits[i] = o;
		}
	}
	if (best_out) memcpy(best_out, c->best, c->nslots * 2);
	return BG_OK;
}
int bg_batch_stats(bg_ctx *c, bg_stats *out) { memset(out, 0, sizeof(*out)); out->tasks = c->ntasks; out->hits = c->nhits; return BG_OK; }

static int finish(bg_ctx *c, int mode, uint16_t *best_inout, bg_hit **hits, uint64_t *nhits) {
	run(c, mode, best_inout);
	bg_hit *h = malloc((c->nhits + 1) * sizeof(bg_hit));
	bg_batch_download(c, h, c->nhits, best_inout);
	*hits = h; *nhits = c->nhits;
	return BG_OK;
}
int bg_align_batch(bg_ctx *c, const bg_queries *Q, const bg_task *tasks, uint64_t ntasks, int mode,
		uint16_t *best_inout, bg_hit **hits, uint64_t *nhits) {
	int rc = bg_batch_upload(c, Q, tasks, ntasks); if (rc) return rc;
	return finish(c, mode, best_inout, hits, nhits);
}
int bg_align_runs(bg_ctx *c, const bg_queries *Q, const bg_run *runs, uint64_t nruns, int mode,
		uint16_t *best_inout, bg_hit **hits, uint64_t *nhits) {
	int rc = bg_batch_upload_runs(c, Q, runs, nruns); if (rc) return rc;
	return finish(c, mode, best_inout, hits, nhits);
}
int bg_align_runs_into(bg_ctx *c, const bg_queries *Q, const bg_run *runs, uint64_t nruns, int mode,
		uint16_t *best_inout, bg_hit *hits, uint64_t cap, uint64_t *nhits) {
	int rc = bg_batch_upload_runs(c, Q, runs, nruns); if (rc) return rc;
	run(c, mode, best_inout);
	*nhits = c->nhits;
	if (c->nhits > cap) return BG_EOVERFLOW;
	return bg_batch_download(c, hits, cap, best_inout);
}
void bg_free_hits(bg_hit *h) { free(h); }

/* compact strand batches: expanded on the host into the general form (strands as byte codes, runs from the bunch lists) */
typedef struct { uint64_t *qoff; uint8_t *codes; uint16_t *bud; uint32_t *slot; } Expanded;
static void expanded_free(Expanded *E) { free(E->qoff); free(E->codes); free(E->bud); free(E->slot); }
static int expand_strands(const bg_reads *R, Expanded *E, const char *who) {
	static const uint8_t RVT[16] = {0, 4, 3, 2, 1, 5, 7, 6, 9, 8, 10, 11, 13, 12, 15, 14};
	memset(E, 0, sizeof(*E));
	uint64_t *roff = malloc(((size_t)R->nreads + 1) * 8), *qoff = malloc(((size_t)R->nq + 1) * 8);
	roff[0] = 0;
	for (uint32_t r = 0; r < R->nreads; ++r) roff[r + 1] = roff[r] + R->len[r];
	qoff[0] = 0;
	for (uint32_t q = 0; q < R->nq; ++q) {
		uint32_t r = R->strand[q] & 0x7FFFFFFFu;
		if (r >= R->nreads) { snprintf(g_err, sizeof(g_err), "%s: strand %u names read %u of %u", who, q, r, R->nreads); free(roff); free(qoff); return BG_EINVAL; }
		qoff[q + 1] = qoff[q] + R->len[r];
	}
	uint8_t *codes = malloc(qoff[R->nq] + 1);
	uint16_t *bud = malloc((size_t)R->nq * 2); uint32_t *slot = malloc((size_t)R->nq * 4);
	for (uint32_t q = 0; q < R->nq; ++q) {
		uint32_t r = R->strand[q] & 0x7FFFFFFFu, len = R->len[r]; int rc = R->strand[q] >> 31;
		bud[q] = R->budget[r]; slot[q] = r;
		for (uint32_t i = 0; i < len; ++i) {
			uint64_t x = roff[r] + (rc ? len - 1 - i : i);
			uint8_t code = R->flags == BG_R_PACKED2 ? (uint8_t)(((R->reads[x >> 2] >> (2 * (x & 3))) & 3) + 1) : (uint8_t)((R->reads[x >> 1] >> (4 * (x & 1))) & 15);
			codes[qoff[q] + i] = rc ? RVT[code] : code;
		}
	}
	free(roff);
	E->qoff = qoff; E->codes = codes; E->bud = bud; E->slot = slot;
	return BG_OK;
}
int bg_align_bunches_into(bg_ctx *c, const bg_reads *R, uint32_t qbunch, const uint32_t *cand_off, const uint32_t *cand, uint32_t nbunch,
		int mode, uint16_t *best_inout, bg_hit *hits, uint64_t cap, uint64_t *nhits) {
	Expanded E;
	int rc = expand_strands(R, &E, "bg_align_bunches_into"); if (rc) return rc;
	uint64_t nruns = cand_off[nbunch];
	bg_run *runs = malloc((nruns + 1) * sizeof(bg_run));
	for (uint32_t b = 0; b < nbunch; ++b) for (uint64_t r = cand_off[b]; r < cand_off[b + 1]; ++r) {
		uint64_t q0 = (uint64_t)b * qbunch;
		runs[r].clump = cand[r]; runs[r].query0 = (uint32_t)q0; runs[r].nq = (uint32_t)(R->nq - q0 < qbunch ? R->nq - q0 : qbunch);
	}
	bg_queries Q = {E.codes, E.qoff, E.bud, E.slot, R->nq, R->nreads, 0};
	rc = bg_align_runs_into(c, &Q, runs, nruns, mode, best_inout, hits, cap, nhits);
	expanded_free(&E); free(runs);
	return rc;
}

/* ---- the accelerator on the "device": the candidate rule of burst.c:4085-4168 restated on the CPU (what k_candgen must reproduce);
 * the order among equal counts is first touch (words ascending, posting order), i.e. a STABLE sort by descending count ---- */
int bg_load_acx(bg_ctx *c, const uint32_t *lens, const uint8_t *postings, uint64_t post_bytes, int word_len, int big, const uint32_t *bad, uint32_t nbad) {
	if (word_len != 12 && word_len != 15) { snprintf(g_err, sizeof(g_err), "bg_load_acx: word length %d (must be 12 or 15)", word_len); return BG_EINVAL; }
	if (!c->num_clumps) { snprintf(g_err, sizeof(g_err), "bg_load_acx: load the database first"); return BG_EINVAL; }
	uint64_t nk = 1ull << (2 * word_len);
	if (c->borrowed) { snprintf(g_err, sizeof(g_err), "bg_load_acx: this context borrows another's database (bg_share_db); load the accelerator there"); return BG_EINVAL; }
	free(c->acx_off); free(c->acx_post); free(c->acx_bad);
	c->acx_off = malloc((nk + 1) * 8); c->acx_off[0] = 0;
	for (uint64_t i = 0; i < nk; ++i) c->acx_off[i + 1] = c->acx_off[i] + (big ? (uint64_t)lens[i] * 3 : (uint64_t)(lens[i] / 2) * 5 + (lens[i] & 1) * 3);
	if (c->acx_off[nk] != post_bytes) { snprintf(g_err, sizeof(g_err), "bg_load_acx: the lengths describe %llu bytes of postings, %llu given", (unsigned long long)c->acx_off[nk], (unsigned long long)post_bytes); return BG_EINVAL; }
	c->acx_post = malloc(post_bytes + 8); memcpy(c->acx_post, postings, post_bytes); memset(c->acx_post + post_bytes, 0, 8);
	c->acx_bad = dup(bad, (size_t)nbad * 4); c->acx_nbad = nbad; c->acx_n = word_len; c->acx_big = big;
	return BG_OK;
}
int bg_share_db(bg_ctx *c, bg_ctx *src) {
	if (!c || !src || c == src) { snprintf(g_err, sizeof(g_err), "bg_share_db: null or identical contexts"); return BG_EINVAL; }
	if (!src->num_clumps) { snprintf(g_err, sizeof(g_err), "bg_share_db: the source context holds no database"); return BG_EINVAL; }
	drop_db(c);
	c->packed = src->packed; c->clump_off = src->clump_off; c->clump_len = src->clump_len; c->num_clumps = src->num_clumps; c->first_clump = src->first_clump;
	c->acx_off = src->acx_off; c->acx_post = src->acx_post; c->acx_bad = src->acx_bad; c->acx_nbad = src->acx_nbad; c->acx_n = src->acx_n; c->acx_big = src->acx_big;
	c->borrowed = 1;
	return BG_OK;
}
static int cmp_u64(const void *a, const void *b) { uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b; return x < y ? -1 : x > y; }
typedef struct { uint32_t clump, count, order; } SimCand;
static int cmp_cand(const void *a, const void *b) { const SimCand *A = a, *B = b; if (A->count != B->count) return A->count > B->count ? -1 : 1; return A->order < B->order ? -1 : A->order > B->order; }
int bg_search_bunches_into(bg_ctx *c, const bg_reads *R, uint32_t qbunch, int heuristic, int skip_bad, int mode,
		uint16_t *best_inout, bg_xhit *hits, uint64_t cap, uint64_t *nhits) {
	if (!c->acx_n) { snprintf(g_err, sizeof(g_err), "bg_search_bunches_into: no accelerator loaded (bg_load_acx)"); return BG_EINVAL; }
	if (R->flags != BG_R_PACKED2) { snprintf(g_err, sizeof(g_err), "bg_search_bunches_into: reads must be BG_R_PACKED2 (plain bases)"); return BG_EINVAL; }
	if (!qbunch || qbunch > BG_RUN_MAX) { snprintf(g_err, sizeof(g_err), "bg_search_bunches_into: bunch size %u", qbunch); return BG_EINVAL; }
	Expanded E;
	int rc = expand_strands(R, &E, "bg_search_bunches_into"); if (rc) return rc;
	const uint32_t N = (uint32_t)c->acx_n, nclumps = c->first_clump + c->num_clumps;
	uint32_t *count = calloc(nclumps, 4), *touched = malloc(((size_t)nclumps + 1) * 4);
	SimCand *cand = malloc(((size_t)nclumps + 1) * sizeof(*cand));
	uint64_t wcap = 1 << 16, *W = malloc(wcap * 8);
	bg_run *runs = NULL; uint64_t nruns = 0, rcap = 0;
	for (uint64_t z = 0; z < R->nq; z += qbunch) {
		const uint32_t nb = (uint32_t)(R->nq - z < qbunch ? R->nq - z : qbunch);
		uint32_t mm[BG_RUN_MAX], minmm = UINT32_MAX; uint64_t nw = 0;
		for (uint32_t j = 0; j < nb; ++j) {
			const uint32_t len = (uint32_t)(E.qoff[z + j + 1] - E.qoff[z + j]), kload = (uint32_t)E.bud[z + j] * N + N;
			uint32_t mmatch = kload < len ? len - kload : 0, heur = heuristic ? (len >> 4) + 1u : 0u;
			if (mmatch < heur) mmatch = heur;
			if (mmatch < minmm) minmm = mmatch;
			mm[j] = kload < len ? len - kload : 1;
			if (nw + len + 1 > wcap) { while (wcap < nw + len + 1) wcap *= 2; W = realloc(W, wcap * 8); }
			const uint8_t *s = E.codes + E.qoff[z + j];
			for (uint32_t k = 0; k + N <= len; ++k) {
				uint64_t w = 0;
				for (uint32_t t = 0; t < N; ++t) w = w << 2 | (uint64_t)(s[k + t] - 1);
				W[nw++] = w << 32 | j;
			}
		}
		qsort(W, nw, 8, cmp_u64);
		uint32_t ntouched = 0;
		for (uint64_t i = 0; i < nw;) {
			const uint32_t v = (uint32_t)(W[i] >> 32); uint32_t mx = 0; uint64_t e = i;
			while (e < nw && (uint32_t)(W[e] >> 32) == v) { uint64_t r = e; while (r < nw && W[r] == W[e]) ++r; if (r - e > mx) mx = (uint32_t)(r - e); e = r; }
			const uint8_t *p = c->acx_post + c->acx_off[v], *end = c->acx_post + c->acx_off[(uint64_t)v + 1];
			while (p < end) {
				uint32_t ids[2], n = 0;
				if (c->acx_big) { ids[n++] = (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16; p += 3; }
				else {
					ids[n++] = ((uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16) & 0xFFFFF;
					if (end - p >= 5) { ids[n++] = ((uint32_t)p[2] >> 4 | (uint32_t)p[3] << 4 | (uint32_t)p[4] << 12) & 0xFFFFF; p += 5; } else p += 3;
				}
				for (uint32_t t = 0; t < n; ++t) if (ids[t] < nclumps) { if (!count[ids[t]]) touched[ntouched++] = ids[t]; count[ids[t]] += mx; }
			}
			i = e;
		}
		uint32_t ncand = 0;
		for (uint32_t i = 0; i < ntouched; ++i) {
			uint32_t v = count[touched[i]]; if (v > 65535) v = 65535; count[touched[i]] = 0;
			if (v > minmm) { cand[ncand].clump = touched[i]; cand[ncand].count = v; cand[ncand].order = ncand; ++ncand; }
		}
		qsort(cand, ncand, sizeof(*cand), cmp_cand);
		const uint64_t need = nruns + (uint64_t)ncand * nb + c->acx_nbad + 1;
		if (need > rcap) { rcap = need * 2; runs = realloc(runs, rcap * sizeof(bg_run)); }
		for (uint32_t i = 0; i < ncand; ++i) {
			uint32_t a = 0;
			while (a < nb) {
				while (a < nb && !(cand[i].count > mm[a])) ++a;
				uint32_t b = a;
				while (b < nb && cand[i].count > mm[b]) ++b;
				if (b > a) { runs[nruns].clump = cand[i].clump; runs[nruns].query0 = (uint32_t)z + a; runs[nruns++].nq = b - a; }
				a = b;
			}
		}
		if (!skip_bad) for (uint32_t i = 0; i < c->acx_nbad; ++i) if (c->acx_bad[i] < nclumps) { runs[nruns].clump = c->acx_bad[i]; runs[nruns].query0 = (uint32_t)z; runs[nruns++].nq = nb; }
	}
	free(count); free(touched); free(cand); free(W);
	bg_queries Q = {E.codes, E.qoff, E.bud, E.slot, R->nq, R->nreads, 0};
	bg_hit *h = NULL; uint64_t n = 0;
	rc = nruns ? bg_align_runs(c, &Q, runs, nruns, mode, best_inout, &h, &n) : BG_OK;
	if (!rc) {
		*nhits = n;
		if (n > cap) { snprintf(g_err, sizeof(g_err), "bg_search_bunches_into: %llu hits, room for %llu", (unsigned long long)n, (unsigned long long)cap); rc = BG_EOVERFLOW; }
		else for (uint64_t i = 0; i < n; ++i) {
			const bg_run *r = runs + (h[i].task >> 4);
			bg_xhit x = {r->query0 + (h[i].task & 15), r->clump, h[i].lane, h[i].ed, h[i].gap_q, h[i].gap_r, h[i].final_pos};
			hits[i] = x;
		}
	}
	bg_free_hits(h); expanded_free(&E); free(runs);
	return rc;
}
