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
	uint8_t *packed; uint64_t *clump_off; uint32_t *clump_len; uint32_t num_clumps, first_clump;
	/* batch */
	uint8_t *codes; uint64_t *qoff; uint16_t *budget; uint32_t *slot; uint32_t nq, nslots;
	uint32_t *tq, *tc; uint64_t *orig; uint64_t ntasks;
	uint16_t *best; OracleHit *hits; uint64_t nhits;
	uint32_t *best32;           /* what bg_batch_best_device() exposes between run_extend and run_select */
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
void bg_free(bg_ctx *c) { if (!c) return; free_batch(c); free(c->packed); free(c->clump_off); free(c->clump_len); free(c); }
void *bg_host_alloc(uint64_t bytes) { return malloc(bytes ? bytes : 1); }
void bg_host_free(void *p) { free(p); }
int bg_set_stream(bg_ctx *c, void *s) { (void)c; (void)s; return BG_OK; }
int bg_set_param(bg_ctx *c, int what, int value) { (void)c; (void)what; (void)value; return BG_OK; }
int bg_set_scoring(bg_ctx *c, const uint8_t S[256]) { memcpy(c->S, S, 256); return BG_OK; }

int bg_load_db(bg_ctx *c, const uint8_t *packed, const uint32_t *clump_len, uint32_t n, uint32_t first) {
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
int bg_align_bunches_into(bg_ctx *c, const bg_reads *R, uint32_t qbunch, const uint32_t *cand_off, const uint32_t *cand, uint32_t nbunch,
		int mode, uint16_t *best_inout, bg_hit *hits, uint64_t cap, uint64_t *nhits) {
	static const uint8_t RVT[16] = {0, 4, 3, 2, 1, 5, 7, 6, 9, 8, 10, 11, 13, 12, 15, 14};
	uint64_t *roff = malloc(((size_t)R->nreads + 1) * 8), *qoff = malloc(((size_t)R->nq + 1) * 8);
	roff[0] = 0;
	for (uint32_t r = 0; r < R->nreads; ++r) roff[r + 1] = roff[r] + R->len[r];
	qoff[0] = 0;
	for (uint32_t q = 0; q < R->nq; ++q) {
		uint32_t r = R->strand[q] & 0x7FFFFFFFu;
		if (r >= R->nreads) { snprintf(g_err, sizeof(g_err), "bg_align_bunches_into: strand %u names read %u of %u", q, r, R->nreads); free(roff); free(qoff); return BG_EINVAL; }
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
	uint64_t nruns = cand_off[nbunch];
	bg_run *runs = malloc((nruns + 1) * sizeof(bg_run));
	for (uint32_t b = 0; b < nbunch; ++b) for (uint64_t r = cand_off[b]; r < cand_off[b + 1]; ++r) {
		uint64_t q0 = (uint64_t)b * qbunch;
		runs[r].clump = cand[r]; runs[r].query0 = (uint32_t)q0; runs[r].nq = (uint32_t)(R->nq - q0 < qbunch ? R->nq - q0 : qbunch);
	}
	bg_queries Q = {codes, qoff, bud, slot, R->nq, R->nreads, 0};
	int rc = bg_align_runs_into(c, &Q, runs, nruns, mode, best_inout, hits, cap, nhits);
	free(roff); free(qoff); free(codes); free(bud); free(slot); free(runs);
	return rc;
}
