/* burst_b200.c -- host driver of the B200 BURST alignment path: BURST's command line, FASTA /
 * .edx / .acx / taxonomy inputs and .b6 output around the CUDA engine of include/burst_b200.h.
 *
 * Written from scratch against the behaviour of knights-lab/BURST burst.c (cited as file:line
 * below); nothing here computes a DP cell -- every (query, clump) pair goes to the GPU through
 * bg_align_batch(), and there is no CPU fallback for it.
 *
 *   CLI + flags            burst.c:4902-5103          -> main()
 *   query preprocessing    burst.c:2980-3223          -> load_queries()
 *   FASTA references       burst.c:1837-1851, 2146-2190, 2687-2741 -> load_fasta_refs()
 *   .edx reader            burst.c:2842-2975          -> load_edx()
 *   .acx reader + lookup   burst.c:3535-3594, 3238-3282, 4085-4133 -> load_acx(), accel_search()
 *   task walk / pods       burst.c:4136-4312, 4343-4519 -> search_all_vs_all(), accel_search()
 *   reporters              burst.c:4553-4891          -> report_*()
 *   taxonomy               burst.c:407-479            -> load_taxonomy(), taxa_lookup_*()
 */
#define _GNU_SOURCE
#define _FILE_OFFSET_BITS 64
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <inttypes.h>
#include <time.h>
#include <pthread.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "burst_b200.h"
#ifdef BURST_NCCL
#include <cuda_runtime_api.h>
#include <nccl.h>
#endif

#define VER "v1.0-b200"
#define VECSZ 16
#define MIN(a, b) ((a) < (b) ? (a) : (b))

typedef enum { FORAGE, BEST, ALLPATHS, CAPITALIST, ANY } Mode;

/* ---- options (defaults of burst.c:81-94, 164) ---- */
static Mode RUNMODE = CAPITALIST;
static float THRES = 0.97f;
static int Z = 1, DO_ACCEL = 0, DO_HEUR = 0, TAXA_NCBI = 0, SCOUR_N = 0;
static uint32_t TAXACUT = 10, LATENCY = 16;
static long REBASE_AMT = 500, DB_QLEN = 500; static int REBASE = 0, ACX_N = 12;   /* ACX_N: word length of the accelerator -d writes (a compile-time constant of the reference binary, burst.c:96-98) */
static float TAXLEVELS_STRICT[] = {.65f, .75f, .78f, .82f, .86f, .94f, .98f, .995f},
             TAXLEVELS_LENIENT[] = {.55f, .70f, .75f, .80f, .84f, .93f, .97f, .985f},   /* burst.c:264-266 */
             *TAXLEVELS = TAXLEVELS_LENIENT;
static int QUIET = 0, GPU_DEVICE = 0, NGPU = 1, THREADS = 1, SHARD_REFS = 0, DEVICE_CAND = 0;

static uint8_t CHAR2NUM[256];
static const uint8_t RVT[16] = {0, 4, 3, 2, 1, 5, 7, 6, 9, 8, 10, 11, 13, 12, 15, 14};   /* burst.c:168 */

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static void *xmalloc(size_t n) { void *p = malloc(n ? n : 1); if (!p) { fputs("OOM\n", stderr); exit(3); } return p; }
static void *xcalloc(size_t n, size_t s) { void *p = calloc(n ? n : 1, s); if (!p) { fputs("OOM\n", stderr); exit(3); } return p; }
static void *xrealloc(void *o, size_t n) { void *p = realloc(o, n ? n : 1); if (!p) { fputs("OOM\n", stderr); exit(3); } return p; }

/* burst.c:1288-1307: letters -> 5 (N) unless named; 'z' and all non-letters -> 0 */
static void init_char2num(void) {
	memset(CHAR2NUM, 0, sizeof(CHAR2NUM));
	for (int c = 'A'; c <= 'Z'; ++c) CHAR2NUM[c] = 5;
	for (int c = 'a'; c < 'z'; ++c) CHAR2NUM[c] = 5;
	const char *L = "ACGTNKMRYSWBVHD";
	for (int i = 0; L[i]; ++i) if (L[i] != 'N') CHAR2NUM[(int)L[i]] = CHAR2NUM[(int)L[i] + 32] = (uint8_t)(i + 1);
	CHAR2NUM['U'] = CHAR2NUM['u'] = 4;
}
static void translate(char *s, size_t n) { for (size_t i = 0; i < n; ++i) s[i] = (char)CHAR2NUM[(uint8_t)s[i]]; }

/* =============================================================================================
 * Queries (burst.c:2980-3223)
 * ============================================================================================= */
typedef struct { char *seq; uint64_t six; uint64_t key; uint8_t rc; } UniBin;      /* burst.c:269-274; key: the first 16 codes, see seq_key() */
typedef struct { uint32_t len; uint16_t ed; } ShrBin;                /* burst.c:277-280 */
typedef struct {
	char **QHead; uint64_t totQ, numUniqQ, newUniqQ, *Offset, QBins[5];
	uint32_t maxLenQ, minLenQ;
	UniBin *UniBins; ShrBin *ShrBins;
	int rc, incl_whitespace, skipAmbig;
} Queries;

/* The reference orders queries by strcmp on their code strings (burst.c:3014-3031).  A sort that follows two pointers per comparison
 * misses the cache twice per comparison; the first 16 codes (1..15, one nibble each, first base in the top nibble, 0 past the end) as
 * one integer order exactly as strcmp orders those 16 bytes, so the sorts compare that key first and only look at the strings when
 * the keys tie (for random 100-base reads: almost never). */
static uint64_t seq_key(const char *s) {
	uint64_t k = 0; int i = 0;
	for (; i < 16 && s[i]; ++i) k = (k << 4) | ((uint8_t)s[i] & 15);
	return i < 16 ? k << (4 * (16 - i)) : k;
}
typedef struct { uint64_t key, ix; } KeyIx;
static char **g_sortseq;
static int cmp_query_ix(const void *a, const void *b) {
	const KeyIx *A = a, *B = b;
	if (A->key != B->key) return A->key < B->key ? -1 : 1;
	int c = strcmp(g_sortseq[A->ix], g_sortseq[B->ix]);
	return c ? c : (A->ix < B->ix ? -1 : A->ix > B->ix);      /* input order among duplicates = the reference at -t 1 */
}
static int cmp_unibin(const void *a, const void *b) {
	const UniBin *A = a, *B = b;
	if (A->key != B->key) return A->key < B->key ? -1 : 1;
	int c = strcmp(A->seq, B->seq);
	if (c) return c;
	if (A->rc != B->rc) return A->rc - B->rc;
	return A->six < B->six ? -1 : A->six > B->six;
}

/* qsort over all host threads: sorted pieces, then rounds of pairwise merges (both comparators here are total orders, so the
 * result does not depend on how the work was cut) */
static void psort(void *base, uint64_t n, size_t sz, int (*cmp)(const void *, const void *)) {
	int T = THREADS < 1 ? 1 : THREADS;
	if (T > 64) T = 64;
	if (T == 1 || n < 65536) { qsort(base, n, sz, cmp); return; }
	int P = 1; while (P * 2 <= T) P *= 2;                              /* pieces: a power of two */
	uint64_t cut[65];
	for (int i = 0; i <= P; ++i) cut[i] = n * (uint64_t)i / (uint64_t)P;
	#pragma omp parallel for schedule(static, 1) num_threads(P)
	for (int i = 0; i < P; ++i) qsort((char *)base + cut[i] * sz, cut[i + 1] - cut[i], sz, cmp);
	char *tmp = xmalloc(n * sz), *src = base, *dst = tmp;
	for (int w = 1; w < P; w *= 2) {
		#pragma omp parallel for schedule(static, 1) num_threads(P / (2 * w))
		for (int i = 0; i < P; i += 2 * w) {
			uint64_t a = cut[i], am = cut[i + w], b = am, bm = cut[i + 2 * w], o = a;
			while (a < am && b < bm) { if (cmp(src + b * sz, src + a * sz) < 0) memcpy(dst + o++ * sz, src + b++ * sz, sz); else memcpy(dst + o++ * sz, src + a++ * sz, sz); }
			if (a < am) memcpy(dst + o * sz, src + a * sz, (am - a) * sz); else if (b < bm) memcpy(dst + o * sz, src + b * sz, (bm - b) * sz);
		}
		char *t = src; src = dst; dst = t;
	}
	if (src != (char *)base) memcpy(base, src, n * sz);
	free(tmp);
}

/* strict two-line FASTA, as parse_tl_faster (burst.c:636-690) */
static void load_queries(const char *fn, Queries *Q) {
	FILE *f = fopen(fn, "rb");
	if (!f) { fprintf(stderr, "Cannot open FASTA file: %s.\n", fn); exit(2); }
	fseeko(f, 0, SEEK_END); uint64_t sz = ftello(f); rewind(f);
	char *dump = xmalloc(sz + 16);
	if (fread(dump, 1, sz, f) != sz) { fputs("ERROR: short read on queries\n", stderr); exit(2); }
	fclose(f);
	memset(dump + sz, 0, 16);
	if (!sz || *dump != '>') { fputs("ERROR: Malformatted FASTA file.\n", stderr); exit(1); }
	uint64_t numNL = 0, numLT = 0;
	#pragma omp parallel for reduction(+:numNL,numLT) schedule(static) num_threads(THREADS < 1 ? 1 : THREADS)
	for (uint64_t i = 0; i < sz; ++i) numNL += dump[i] == '\n', numLT += dump[i] == '>';
	numNL += numNL & 1;
	if (numLT != numNL / 2) { fputs("ERROR: line count != '>' * 2\n", stderr); exit(1); }
	uint64_t totQ = numLT;
	char **Head = xmalloc(totQ * sizeof(*Head)), **Seq = xmalloc(totQ * sizeof(*Seq));
	uint32_t *Len = xmalloc(totQ * sizeof(*Len));
	uint64_t n = 0; char *p = dump;
	while (p < dump + sz && *p == '>') {
		char *h = p + 1, *nl = memchr(h, '\n', dump + sz - h);
		if (!nl) break;
		*nl = 0; if (nl > h && nl[-1] == '\r') nl[-1] = 0;
		char *s = nl + 1, *e = memchr(s, '\n', dump + sz - s);
		if (!e) e = dump + sz;
		*e = 0; if (e > s && e[-1] == '\r') *--e = 0;
		Head[n] = h; Seq[n] = s; Len[n] = (uint32_t)(e - s); ++n;
		p = (e < dump + sz) ? e + 1 : e;
		while (p < dump + sz && *p != '>') ++p;       /* tolerate the \r\n tail */
	}
	totQ = n;
	if (!totQ) { fputs("ERROR: No queries found.", stderr); exit(1); }
	if (!Q->incl_whitespace) for (uint64_t i = 0; i < totQ; ++i) {    /* burst.c:2987-2993 */
		char *q = Head[i]; while (*q && *q != ' ' && *q != '\t') ++q; *q = 0;
	}
	uint32_t maxLenQ = 0, minLenQ = UINT32_MAX;
	for (uint64_t i = 0; i < totQ; ++i) { if (Len[i] > maxLenQ) maxLenQ = Len[i]; if (Len[i] < minLenQ) minLenQ = Len[i]; }
	if (maxLenQ > (1 << 16)) fputs("WARNING: Max query length is very long\n", stderr);
	if (minLenQ < 5) fputs("WARNING: Min query length is less than 5 bases\n", stderr);
	printf("Parsed %" PRIu64 " queries. Found min %u, max %u.\n", totQ, minLenQ, maxLenQ);
	#pragma omp parallel for schedule(static, 4096) num_threads(THREADS < 1 ? 1 : THREADS)
	for (uint64_t i = 0; i < totQ; ++i) translate(Seq[i], Len[i]);
	/* sort (burst.c:3014-3031): strcmp on the code strings */
	KeyIx *kx = xmalloc(totQ * sizeof(*kx));
	#pragma omp parallel for schedule(static, 4096) num_threads(THREADS < 1 ? 1 : THREADS)
	for (uint64_t i = 0; i < totQ; ++i) { kx[i].key = seq_key(Seq[i]); kx[i].ix = i; }
	g_sortseq = Seq;
	psort(kx, totQ, sizeof(*kx), cmp_query_ix);
	/* uniqueness (burst.c:3036-3053): a new sequence starts where the key or, on equal keys, the string changes */
	uint64_t *ix = xmalloc(totQ * sizeof(*ix));
	uint8_t *fresh = xmalloc(totQ);
	uint64_t numUniq = 0;
	#pragma omp parallel for reduction(+:numUniq) schedule(static, 4096) num_threads(THREADS < 1 ? 1 : THREADS)
	for (uint64_t i = 0; i < totQ; ++i) {
		ix[i] = kx[i].ix;
		fresh[i] = !i || kx[i - 1].key != kx[i].key || strcmp(Seq[kx[i - 1].ix], Seq[kx[i].ix]) != 0;
		numUniq += fresh[i];
	}
	free(kx);
	uint64_t *Offset = xmalloc((numUniq + 1) * sizeof(*Offset)), u = 0;
	for (uint64_t i = 0; i < totQ; ++i) if (fresh[i]) Offset[u++] = i;
	Offset[numUniq] = totQ;
	free(fresh);
	char **SrtHead = xmalloc(totQ * sizeof(*SrtHead));
	#pragma omp parallel for schedule(static, 4096) num_threads(THREADS < 1 ? 1 : THREADS)
	for (uint64_t i = 0; i < totQ; ++i) SrtHead[i] = Head[ix[i]];
	uint64_t newUniq = numUniq * (Q->rc ? 2 : 1);
	UniBin *UB = xcalloc(newUniq, sizeof(*UB)); ShrBin *SB = xcalloc(numUniq, sizeof(*SB));
	float reqID = 1 / THRES - 1;                                      /* burst.c:3069-3076, float32 */
	uint64_t uniqTotLen = 0;
	#pragma omp parallel for reduction(+:uniqTotLen) schedule(static, 4096) num_threads(THREADS < 1 ? 1 : THREADS)
	for (uint64_t i = 0; i < numUniq; ++i) {
		uint64_t r = ix[Offset[i]];
		uint32_t len = Len[r], ed = reqID * len;
		SB[i].len = len; SB[i].ed = (uint16_t)MIN(254, ed);
		UB[i].seq = Seq[r]; UB[i].six = i; UB[i].rc = 0; UB[i].key = seq_key(Seq[r]);
		uniqTotLen += len;
	}
	if (Q->rc) {                                                      /* burst.c:3087-3109 */
		char *rcd = xmalloc(uniqTotLen + numUniq + 1);
		uint64_t *woff = xmalloc((numUniq + 1) * sizeof(*woff));
		woff[0] = 0;
		for (uint64_t i = 0; i < numUniq; ++i) woff[i + 1] = woff[i] + SB[i].len + 1;
		#pragma omp parallel for schedule(static, 4096) num_threads(THREADS < 1 ? 1 : THREADS)
		for (uint64_t i = 0; i < numUniq; ++i) {
			uint32_t len = SB[i].len; char *org = UB[i].seq, *w = rcd + woff[i];
			for (uint32_t j = 0; j < len; ++j) w[j] = (char)RVT[(uint8_t)org[len - j - 1] & 15];
			w[len] = 0;
			UB[numUniq + i].seq = w; UB[numUniq + i].rc = 1; UB[numUniq + i].six = i; UB[numUniq + i].key = seq_key(w);
		}
		free(woff);
	}
	memset(Q->QBins, 0, sizeof(Q->QBins));
	if (DO_ACCEL) {                                                   /* burst.c:3113-3177 */
		uint8_t *stat = xmalloc(newUniq + 1);
		#pragma omp parallel for schedule(static, 4096) num_threads(THREADS < 1 ? 1 : THREADS)
		for (uint64_t i = 0; i < newUniq; ++i) {
			uint32_t len = SB[UB[i].six].len, ed = SB[UB[i].six].ed, totN = 0;
			const char *s = UB[i].seq; stat[i] = 1;
			if (len < (uint32_t)SCOUR_N || (!DO_HEUR && ed >= len / (uint32_t)SCOUR_N)) stat[i] = 2;
			else for (uint32_t j = 0; j < len; ++j) {
				if ((totN += s[j] > 4 + Z) > 5) { stat[i] = 2; break; }
				else if (s[j] > 4) stat[i] = 0;
			}
		}
		/* stable partition into ambiguous (0), clear (1), bad (2); each bin is re-sorted below */
		UniBin *T = xmalloc(newUniq * sizeof(*T)); uint64_t c[3] = {0, 0, 0}, o[3];
		for (uint64_t i = 0; i < newUniq; ++i) ++c[stat[i]];
		o[0] = 0; o[1] = c[0]; o[2] = c[0] + c[1];
		Q->QBins[0] = c[0]; Q->QBins[1] = c[0] + c[1]; Q->QBins[2] = newUniq;
		for (uint64_t i = 0; i < newUniq; ++i) T[o[stat[i]]++] = UB[i];
		memcpy(UB, T, newUniq * sizeof(*T)); free(T); free(stat);
		printf("Unambig: %" PRIu64 ", ambig: %" PRIu64 ", super-ambig: %" PRIu64 "\n", c[1], c[0], c[2]);
		psort(UB, Q->QBins[0], sizeof(*UB), cmp_unibin);
		psort(UB + Q->QBins[0], Q->QBins[1] - Q->QBins[0], sizeof(*UB), cmp_unibin);
		psort(UB + Q->QBins[1], Q->QBins[2] - Q->QBins[1], sizeof(*UB), cmp_unibin);
	} else if (Q->rc) psort(UB, newUniq, sizeof(*UB), cmp_unibin);   /* burst.c:3178-3186 */
	free(ix); free(Head); free(Seq); free(Len);
	Q->QHead = SrtHead; Q->totQ = totQ; Q->numUniqQ = numUniq; Q->newUniqQ = newUniq; Q->Offset = Offset;
	Q->UniBins = UB; Q->ShrBins = SB; Q->maxLenQ = maxLenQ; Q->minLenQ = minLenQ;
	printf("Number unique: %" PRIu64 "\n", numUniq);
}

/* =============================================================================================
 * References
 * ============================================================================================= */
typedef struct {
	char **RefHead; uint32_t *ClumpLen, *RefStart, *RefIxSrt, *TmpRIX, *RefDedupIx, *RefMap;
	uint32_t totR, origTotR, numRclumps, maxLenR, numRefHeads, shear;
	uint8_t *packed;                 /* clumps in .edx layout (burst.c:2810-2824) */
	uint64_t packedBytes;
} Refs;

/* multi-line FASTA reader in the manner of parse_tl_fasta (burst.c:484-535) */
static uint32_t read_fasta_refs(const char *fn, char ***HeadP, char ***SeqP, uint32_t **LenP) {
	FILE *f = fopen(fn, "rb");
	if (!f) { fprintf(stderr, "Cannot open FASTA file: %s.\n", fn); exit(2); }
	size_t cap = 1024, ns = 0; int lastHd = 0, have = 0;
	char **Head = xmalloc(cap * sizeof(*Head)), **Seq = xmalloc(cap * sizeof(*Seq));
	uint32_t *Len = xmalloc(cap * sizeof(*Len));
	char *line = NULL; size_t lcap = 0; ssize_t got;
	while ((got = getline(&line, &lcap, f)) >= 0) {
		size_t len = (size_t)got;
		if (len && line[len - 1] == '\n') --len;
		if (len && line[len - 1] == '\r') --len;
		line[len] = 0;
		if (*line == '>') {
			if (lastHd) continue;                      /* a header right after a header is ignored */
			if (have) ++ns;
			if (ns == cap) { cap *= 2; Head = xrealloc(Head, cap * sizeof(*Head)); Seq = xrealloc(Seq, cap * sizeof(*Seq)); Len = xrealloc(Len, cap * sizeof(*Len)); }
			have = 1; lastHd = 1;
			Head[ns] = strdup(line + 1); Len[ns] = 0; Seq[ns] = NULL;
		} else if (*line == 0 || *line == ' ') continue;
		else {
			if (!have) continue;
			lastHd = 0;
			Seq[ns] = xrealloc(Seq[ns], (size_t)Len[ns] + len + 17);
			memcpy(Seq[ns] + Len[ns], line, len);
			Len[ns] += (uint32_t)len;
			memset(Seq[ns] + Len[ns], 0, 17);
		}
	}
	free(line); fclose(f);
	if (have) { if (lastHd) puts("WARNING: file ends on header. Skipping last sequence."); else ++ns; }
	*HeadP = Head; *SeqP = Seq; *LenP = Len;
	return (uint32_t)ns;
}

typedef struct { const char *seq; uint32_t len, ix; } Tux;
static int cmp_tux_len(const void *a, const void *b) {
	const Tux *A = a, *B = b;
	if (A->len != B->len) return A->len < B->len ? -1 : 1;
	return A->ix < B->ix ? -1 : A->ix > B->ix;
}
static int cmp_tux_seq(const void *a, const void *b) {
	const Tux *A = a, *B = b;
	uint32_t ml = MIN(A->len, B->len);
	int c = memcmp(A->seq, B->seq, ml);
	if (c) return c;
	if (A->len != B->len) return A->len < B->len ? -1 : 1;
	return A->ix < B->ix ? -1 : A->ix > B->ix;
}

/* Pack code strings 16 per clump in .edx clump layout; lanes beyond a reference's end hold 0. */
static void pack_clumps(Refs *R, char **Seq, const uint32_t *Len) {
	uint32_t totR = R->totR, nfull = totR / VECSZ, totRC = nfull + (nfull * VECSZ < totR);
	R->numRclumps = totRC;
	R->ClumpLen = xcalloc(totRC + 1, sizeof(*R->ClumpLen));
	uint64_t bytes = 0;
	for (uint32_t c = 0; c < totRC; ++c) {
		uint32_t cl = 0;
		for (uint32_t k = 0; k < VECSZ && c * VECSZ + k < totR; ++k) { uint32_t l = Len[R->RefIxSrt[c * VECSZ + k]]; if (l > cl) cl = l; }
		if (!cl) cl = 1;
		R->ClumpLen[c] = cl; bytes += (uint64_t)((cl + 1) / 2) * 16;
	}
	R->packed = xcalloc(bytes + 16, 1); R->packedBytes = bytes;
	uint8_t *w = R->packed;
	for (uint32_t c = 0; c < totRC; ++c) {
		uint32_t cl = R->ClumpLen[c];
		for (uint32_t k = 0; k < VECSZ && c * VECSZ + k < totR; ++k) {
			uint32_t r = R->RefIxSrt[c * VECSZ + k], l = Len[r]; const char *s = Seq[r];
			for (uint32_t j = 0; j < l; ++j) w[(size_t)(j >> 1) * 16 + k] |= (uint8_t)((s[j] & 15) << ((j & 1) * 4));
		}
		w += (size_t)((cl + 1) / 2) * 16;
	}
}

/* plain -r FASTA (burst.c:1837-1851 parse/translate, 2146-2190 ordering, 2687-2741 packing) */
static void load_fasta_refs(const char *fn, Refs *R) {
	char **Head, **Seq; uint32_t *Len;
	uint32_t totR = read_fasta_refs(fn, &Head, &Seq, &Len);
	printf("Parsed %u references.\n", totR);
	if (!totR) { fputs("ERROR: no references found.\n", stderr); exit(1); }
	for (uint32_t i = 0; i < totR; ++i) { if (!Seq[i]) Seq[i] = xcalloc(17, 1); translate(Seq[i], Len[i]); }
	if (REBASE) fputs("WARNING: -s on a FASTA reference is not supported by this build; references are used whole.\n", stderr);
	memset(R, 0, sizeof(*R));
	R->totR = R->origTotR = totR; R->RefHead = Head;
	/* references of similar length (within LATENCY bases) share clumps, lexicographic inside a pod */
	Tux *T = xmalloc(totR * sizeof(*T));
	for (uint32_t i = 0; i < totR; ++i) T[i] = (Tux){Seq[i], Len[i], i};
	qsort(T, totR, sizeof(*T), cmp_tux_len);
	R->maxLenR = T[totR - 1].len;
	uint32_t prev = 0, curTol = T[0].len;
	for (uint32_t i = 1; i <= totR; ++i) if (i == totR || T[i].len > curTol + LATENCY) {
		if (i - prev > 1) qsort(T + prev, i - prev, sizeof(*T), cmp_tux_seq);
		if (i < totR) curTol = T[i].len;
		prev = i;
	}
	R->RefIxSrt = xmalloc((totR + 1) * sizeof(*R->RefIxSrt));
	for (uint32_t i = 0; i < totR; ++i) R->RefIxSrt[i] = T[i].ix;
	free(T);
	R->TmpRIX = R->RefIxSrt;
	pack_clumps(R, Seq, Len);
	printf("There are %u references and hence %u clumps\n", totR, R->numRclumps);
	for (uint32_t i = 0; i < totR; ++i) free(Seq[i]);
	free(Seq); free(Len);
}

static void rd(void *dst, size_t sz, size_t n, FILE *f) {
	if (fread(dst, sz, n, f) != n) { fputs("ERROR: truncated database file\n", stderr); exit(1); }
}
/* a large section of a file read by all host threads at once (page faults and copies of a 4 GB table are the cost, not the disk) */
static void rd_at(FILE *f, void *dst, uint64_t bytes, uint64_t off) {
	const int fd = fileno(f); const uint64_t CH = 32ull << 20; const int64_t nch = (int64_t)((bytes + CH - 1) / CH); int bad = 0;
	#pragma omp parallel for schedule(dynamic, 1) num_threads(THREADS < 1 ? 1 : THREADS)
	for (int64_t c = 0; c < nch; ++c) {
		uint64_t a = (uint64_t)c * CH, n = MIN(CH, bytes - a);
		while (n) { ssize_t g = pread(fd, (char *)dst + a, n, (off_t)(off + a)); if (g <= 0) { bad = 1; break; } a += (uint64_t)g; n -= (uint64_t)g; }
	}
	if (bad) { fputs("ERROR: truncated database file\n", stderr); exit(1); }
}
/* .edx reader (burst.c:2842-2975; layout written at 2758-2839) */
static void load_edx(const char *fn, Refs *R) {
	FILE *in = fopen(fn, "rb");
	if (!in) { fputs("ERROR: cannot parse EDB", stderr); exit(1); }
	memset(R, 0, sizeof(*R));
	uint8_t cb = (uint8_t)fgetc(in), ver = cb & 0xF;
	if (ver != 3 && ver != 2) { fprintf(stderr, "ERROR: invalid database version %u\n", ver); exit(1); }
	if (ver == 2) { fprintf(stderr, "ERROR: Old DB version. Re-make with new version.\n"); exit(2); }
	REBASE = (cb >> 6) & 1;
	printf(" --> EDB: Fingerprints are DISABLED\n");                     /* burst.c:2856-2857: DO_FP = (DB has them) && -f; -f is not supported here, so always */
	if ((cb >> 4) & 1) { fprintf(stderr, "ERROR: DB made with Xalpha; queries must use Xalpha.\n"); exit(1); }
	uint64_t totRefHeadLen; uint32_t shear, totR, origTotR, numRclumps, maxLenR, numRefHeads;
	rd(&totRefHeadLen, 8, 1, in); rd(&shear, 4, 1, in); rd(&totR, 4, 1, in); rd(&origTotR, 4, 1, in);
	rd(&numRclumps, 4, 1, in); rd(&maxLenR, 4, 1, in);
	char *hd = xmalloc(totRefHeadLen + 1);
	rd(hd, 1, totRefHeadLen, in);
	rd(&numRefHeads, 4, 1, in);
	char **uniq = xmalloc((size_t)numRefHeads * sizeof(*uniq));
	uniq[0] = hd;
	for (uint32_t i = 1; i < numRefHeads; ++i) { while (*hd++); uniq[i] = hd; }
	R->RefMap = xmalloc((size_t)origTotR * 4);
	rd(R->RefMap, 4, origTotR, in);
	R->RefHead = xmalloc((size_t)origTotR * sizeof(*R->RefHead));
	for (uint32_t i = 0; i < origTotR; ++i) R->RefHead[i] = uniq[R->RefMap[i]];
	free(uniq);
	if (REBASE) { printf(" --> EDB: Sheared database (shear size = %u)\n", shear); R->RefStart = xmalloc((size_t)origTotR * 4); rd(R->RefStart, 4, origTotR, in); }
	if (totR != origTotR) { puts(" --> EDB: Unique-reference database"); R->RefDedupIx = xmalloc(((size_t)totR + 1) * 4); rd(R->RefDedupIx, 4, (size_t)totR + 1, in); }
	R->TmpRIX = xmalloc((size_t)origTotR * 4); rd(R->TmpRIX, 4, origTotR, in);
	R->ClumpLen = xmalloc((size_t)numRclumps * 4); rd(R->ClumpLen, 4, numRclumps, in);
	uint64_t vecs = 0;
	for (uint32_t i = 0; i < numRclumps; ++i) { if (R->ClumpLen[i] > maxLenR) maxLenR = R->ClumpLen[i]; vecs += R->ClumpLen[i] / 2u + (R->ClumpLen[i] & 1); }
	R->packed = xmalloc(vecs * 16 + 16); R->packedBytes = vecs * 16;
	rd(R->packed, 16, vecs, in);
	fclose(in);
	R->totR = totR; R->origTotR = origTotR; R->numRclumps = numRclumps; R->maxLenR = maxLenR;
	R->numRefHeads = numRefHeads; R->shear = shear;
	if (R->RefDedupIx) {                                         /* burst.c:3688-3693 */
		R->RefIxSrt = xmalloc((size_t)totR * 4);
		for (uint32_t i = 0; i < totR; ++i) R->RefIxSrt[i] = R->TmpRIX[R->RefDedupIx[i]];
	} else R->RefIxSrt = R->TmpRIX;
	printf(" --> EDB: %u refs [%u orig], %u clumps, %u maxR\n", totR, origTotR, numRclumps, maxLenR);
}

/* =============================================================================================
 * Taxonomy (burst.c:407-479)
 * ============================================================================================= */
typedef struct { char *Head, *Tax; } TaxPair;
static TaxPair *Taxonomy; static size_t taxa_parsed;
static char NULLTAX[1] = {0};
static int cmp_tax(const void *a, const void *b) { return strcmp(((const TaxPair *)a)->Head, ((const TaxPair *)b)->Head); }
static void load_taxonomy(const char *fn) {
	FILE *f = fopen(fn, "rb");
	if (!f) { fprintf(stderr, "Cannot open TAXONOMY file: %s.\n", fn); exit(2); }
	size_t cap = 1024, ns = 0; TaxPair *T = xmalloc(cap * sizeof(*T));
	char *line = NULL; size_t lcap = 0; ssize_t got;
	while ((got = getline(&line, &lcap, f)) >= 0) {
		if (ns == cap) T = xrealloc(T, (cap *= 2) * sizeof(*T));
		size_t i = 0, j;
		for (; line[i] != '\t'; ++i) if (!line[i]) { fprintf(stderr, "ERROR: invalid taxonomy [%zu]\n", ns); exit(2); }
		T[ns].Head = strndup(line, i);
		for (j = ++i; line[j] && line[j] != '\n' && line[j] != '\r' && line[j] != '\t'; ++j);
		T[ns].Tax = strndup(line + i, j - i);
		++ns;
	}
	free(line); fclose(f);
	if (!ns) { fputs("ERROR: invalid taxonomy\n", stderr); exit(1); }
	qsort(T, ns, sizeof(*T), cmp_tax);
	Taxonomy = T; taxa_parsed = ns;
}
/* exact header match; with -bn the key skips its first 4 characters and may end at a '.' (burst.c:424-440) */
static char *taxa_lookup(const char *key) {
	if (TAXA_NCBI) key += strlen(key) >= 4 ? 4 : strlen(key);
	size_t lo = 0, hi = taxa_parsed;
	while (lo < hi) {
		size_t mid = (lo + hi) / 2;
		const char *r = Taxonomy[mid].Head, *k = key;
		while (*r && *r == *k) ++r, ++k;
		if (!*r && (!*k || (TAXA_NCBI && *k == '.'))) return Taxonomy[mid].Tax;
		unsigned char kc = (unsigned char)((TAXA_NCBI && *k == '.') ? 0 : *k);
		if ((unsigned char)*r < kc) lo = mid + 1; else hi = mid;
	}
	return NULLTAX;
}

/* =============================================================================================
 * Pods: what the reference keeps per unique query (ResultPod, burst.c:3998-4004)
 * ============================================================================================= */
typedef struct { float score; uint32_t refIx, finalPos; uint8_t numGapR, numGapQ, mismatches, rc; } Pod;
typedef struct { Pod *p; uint32_t n, cap; } PodList;

static void pod_push(PodList *L, Pod x) {
	if (L->n == L->cap) { L->cap = L->cap ? L->cap * 2 : 4; L->p = xrealloc(L->p, L->cap * sizeof(Pod)); }
	L->p[L->n++] = x;
}
/* identity exactly as burst.c:844-860: IEEE single divide then subtract */
static float identity(uint32_t ed, uint32_t qlen, uint32_t gapq) {
	volatile float den = (float)qlen + (float)gapq, q = (float)ed / den;
	return 1.0f - q;
}
static void die_gpu(const char *what, int rc) {
	fprintf(stderr, "ERROR: GPU engine failed in %s: %s\n", what, bg_last_error());
	exit(rc == BG_ENOMEM ? 3 : 4);
}


/* =============================================================================================
 * Accelerator (.acx): reader (burst.c:3535-3594) and candidate generation (burst.c:3232-3282,
 * 4085-4133).  The file does not record N (12 or 15, a compile-time constant of the reference
 * binary, burst.c:96-98); it is inferred from the file size.
 * ============================================================================================= */
typedef struct { uint64_t *off; uint8_t *post; uint32_t *bad, nbad; int big; uint64_t nk; uint32_t *lens; uint64_t post_bytes; } Acx;
static const uint8_t AMBIG_N[16] = {0, 1, 1, 1, 1, 4, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3};
static const uint8_t AMBIG_B[16][4] = {{0}, {0}, {1}, {2}, {3}, {0, 1, 2, 3}, {2, 3}, {0, 1}, {0, 2}, {1, 3}, {1, 2}, {0, 3},
	{1, 2, 3}, {0, 1, 2}, {0, 1, 3}, {0, 2, 3}};                              /* burst.c:1372-1375 */

static void load_acx(const char *fn, Acx *A) {
	FILE *in = fopen(fn, "rb");
	if (!in) { fprintf(stderr, "Cannot read accelerator '%s'\n", fn); exit(1); }
	fseeko(in, 0, SEEK_END); uint64_t fsz = ftello(in); rewind(in);
	uint8_t cb = (uint8_t)fgetc(in), ver = cb & 0xF, didZ = (cb >> 6) & 1;
	if (didZ && !Z) { fprintf(stderr, "ERROR: Accelerator built without '-y'; can't use '-y'\n"); exit(1); }
	if (cb < 128 || (ver != 0 && ver != 1)) { fprintf(stderr, "ERROR: invalid accelerator [%u:%u]\n", cb, ver); exit(1); }
	uint32_t szBL; rd(&szBL, 4, 1, in);
	memset(A, 0, sizeof(*A)); A->big = ver == 1;
	int found = 0;
	for (int n = 12; n <= 15 && !found; n += 3) {
		uint64_t nk = 1ull << (2 * n);
		if (fsz < 5 + nk * 4) continue;
		uint32_t *Lens = xmalloc(nk * 4);
		rd_at(in, Lens, nk * 4, 5);
		uint64_t bytes = 0;
		#pragma omp parallel for reduction(+:bytes) schedule(static)
		for (uint64_t i = 0; i < nk; ++i) bytes += A->big ? (uint64_t)Lens[i] * 3 : (uint64_t)(Lens[i] / 2u) * 5 + (Lens[i] & 1) * 3;
		if (5 + nk * 4 + bytes + (uint64_t)szBL * 4 == fsz) {
			found = 1; SCOUR_N = n; A->nk = nk;
			A->off = xmalloc((nk + 1) * 8); A->off[0] = 0;
			{	/* prefix sums in parallel: per-block totals, then each block from its base */
				enum { NB = 256 }; uint64_t base[NB + 1]; const uint64_t per = (nk + NB - 1) / NB;
				#pragma omp parallel for schedule(static)
				for (int b = 0; b < NB; ++b) {
					uint64_t t = 0, lo = (uint64_t)b * per, hi = MIN(nk, lo + per);
					for (uint64_t i = lo; i < hi; ++i) t += A->big ? (uint64_t)Lens[i] * 3 : (uint64_t)(Lens[i] / 2u) * 5 + (Lens[i] & 1) * 3;
					base[b + 1] = t;
				}
				base[0] = 0;
				for (int b = 0; b < NB; ++b) base[b + 1] += base[b];
				#pragma omp parallel for schedule(static)
				for (int b = 0; b < NB; ++b) {
					uint64_t t = base[b], lo = (uint64_t)b * per, hi = MIN(nk, lo + per);
					for (uint64_t i = lo; i < hi; ++i) { t += A->big ? (uint64_t)Lens[i] * 3 : (uint64_t)(Lens[i] / 2u) * 5 + (Lens[i] & 1) * 3; A->off[i + 1] = t; }
				}
			}
			A->post = xmalloc(bytes + 16); rd_at(in, A->post, bytes, 5 + nk * 4); memset(A->post + bytes, 0, 16);
			A->bad = xmalloc((uint64_t)szBL * 4 + 4); fseeko(in, (off_t)(5 + nk * 4 + bytes), SEEK_SET); rd(A->bad, 4, szBL, in); A->nbad = szBL;
			A->post_bytes = bytes;
			if (DEVICE_CAND) { A->lens = Lens; Lens = NULL; }                  /* bg_load_acx takes the on-disk form */
		}
		free(Lens);
	}
	fclose(in);
	if (!found) { fputs("ERROR: accelerator size matches neither a DB12 nor a DB15 layout\n", stderr); exit(1); }
	printf(" --> [Accel] Accelerator found (DB%d, %s format), %u ambiguous entries\n", SCOUR_N, A->big ? "LARGE" : "SMALL", szBL);
}

typedef struct { uint32_t v, i; } Split;
static int cmp_u64(const void *a, const void *b) { uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b; return x < y ? -1 : x > y; }
static int cmp_refcount(const void *a, const void *b) { const Split *A = a, *B = b; return A->i > B->i ? -1 : B->i > A->i; }   /* burst.c:4028-4031 */

static void ambig_words(uint64_t *W, uint64_t *wix, const char *s, uint32_t j, uint32_t w, int ix) {      /* burst.c:3232-3236 */
	if (ix == SCOUR_N) W[(*wix)++] = (uint64_t)w << 32 | j;
	else for (int i = 0; i < AMBIG_N[(uint8_t)s[ix] & 15]; ++i) ambig_words(W, wix, s, j, w << 2 | AMBIG_B[(uint8_t)s[ix] & 15][i], ix + 1);
}

/* =============================================================================================
 * Database creation (-d): FASTA -> .edx (+ .acx with -a).  SURVEY.md 8(f) #3.
 * The byte layouts are the reference's (dump_edb burst.c:2758-2839, make_accelerator 3501-3530), so
 * either binary loads the files.  The way references are cut and grouped is NOT the reference's
 * duplicate-guided shearing / clustering (burst.c:1852-2686): it is the plain shear of its QUICK
 * branch (burst.c:2115-2143: windows of shear + ov bases every `shear` bases, ov = qLen / id) followed
 * by the length / lexicographic ordering of load_fasta_refs.  Clump composition changes the amount
 * of work, never the alignments found.
 * ============================================================================================= */
static int cmp_strp(const void *a, const void *b) { return strcmp(*(char *const *)a, *(char *const *)b); }

typedef struct { uint8_t *yn; uint32_t *cache, n; uint64_t cap; int N; uint32_t mask; } WordSet;   /* distinct words of one clump */
static void ws_add(WordSet *S, uint32_t w) {
	if (S->yn[w >> 3] & (1u << (w & 7))) return;
	S->yn[w >> 3] |= (uint8_t)(1u << (w & 7));
	if (S->n == S->cap) { S->cap *= 2; S->cache = xrealloc(S->cache, S->cap * 4); }
	S->cache[S->n++] = w;
}
static void ws_ambig(WordSet *S, const char *s, uint32_t w, int ix) {      /* every plain variant of an ambiguous window (burst.c:3285-3292) */
	if (ix == S->N) ws_add(S, w);
	else for (int i = 0; i < AMBIG_N[(uint8_t)s[ix] & 15]; ++i) ws_ambig(S, s, w << 2 | AMBIG_B[(uint8_t)s[ix] & 15][i], ix + 1);
}
/* words of one clump's lanes into S; returns 1 if the clump is "bad" (too many ambiguous variants: always visited instead, burst.c:3349-3356) */
static int clump_words(WordSet *S, char **Seq, const uint32_t *Len, const uint32_t *ix, uint32_t nlanes, int skipAmbig) {
	const int N = S->N; const uint32_t AMBIG = 4 + (uint32_t)Z;
	const uint64_t fullSize = N > 14 ? INT32_MAX : (1u << 24);
	S->n = 0;
	uint64_t Tsum = 0; uint16_t doAmbig = 0;
	if (!skipAmbig) for (uint32_t z = 0; z < nlanes; ++z) {
		uint32_t len = Len[ix[z]], Asum = 0; const char *s = Seq[ix[z]];
		if (len < (uint32_t)N) continue;
		for (uint32_t j = 0; j < len; ++j) {
			if (j >= (uint32_t)N - 1) { uint64_t v = 1; for (uint32_t a = 0; a < Asum && v < fullSize; ++a) v *= Z ? 3 : 4; Tsum += v; if ((uint8_t)s[j - (N - 1)] > AMBIG) --Asum; }
			if ((uint8_t)s[j] > AMBIG) ++Asum, doAmbig |= (uint16_t)(1u << z);
			if (Tsum >= fullSize) return 1;
		}
	}
	for (uint32_t z = 0; z < nlanes; ++z) {
		uint32_t len = Len[ix[z]]; const char *s = Seq[ix[z]];
		if (len < (uint32_t)N) continue;
		if (skipAmbig || Z) {                                            /* windows holding an N (any ambiguous base with -sa) are not indexed */
			for (uint32_t j = 0; j + N <= len; ++j) {
				int k = 0;
				for (; k < N; ++k) if (skipAmbig ? (uint8_t)s[j + k] >= 5 : s[j + k] == 5) break;
				if (k < N) { j += k; continue; }
				ws_ambig(S, s + j, 0, 0);
			}
		} else if (doAmbig >> z & 1) for (uint32_t j = 0; j + N <= len; ++j) ws_ambig(S, s + j, 0, 0);
		else {
			uint32_t w = 0;
			for (uint32_t j = 0; j < len; ++j) { w = (w << 2 | (uint32_t)(s[j] - 1)) & S->mask; if (j + 1 >= (uint32_t)N) ws_add(S, w); }
		}
	}
	return 0;
}

static void make_db(const char *ref_FN, const char *edx_FN, const char *acx_FN, long dbQLen, int N, int skipAmbig) {
	char **Head, **Seq; uint32_t *Len;
	uint32_t origR = read_fasta_refs(ref_FN, &Head, &Seq, &Len);
	printf("Parsed %u references.\n", origR);
	if (!origR) { fputs("ERROR: no references found.\n", stderr); exit(1); }
	for (uint32_t i = 0; i < origR; ++i) { if (!Seq[i]) Seq[i] = xcalloc(17, 1); translate(Seq[i], Len[i]); }
	if (!REBASE) dbQLen = 0;                                              /* burst.c:5121 */
	/* ---- shear (burst.c:1853-1856, 2115-2143) ---- */
	uint32_t totR = origR, *Start = NULL, *Src = NULL, maxLenR = 0;
	char **SSeq = Seq; uint32_t *SLen = Len;
	if (REBASE) {
		uint32_t minShear = (uint32_t)(dbQLen / THRES), shear = minShear > (uint32_t)REBASE_AMT ? minShear : (uint32_t)REBASE_AMT, ov = minShear;
		printf("\nInitiating database shearing procedure [shear %u, ov %u].\n", shear, ov);
		uint64_t n = 0;
		for (uint32_t i = 0; i < origR; ++i) { long unit = (long)Len[i] - (long)ov; if (unit <= 0) unit = 1; n += (uint64_t)(unit / shear + (unit % shear != 0)); }
		if (n >= UINT32_MAX) { fputs("ERROR: too many shears\n", stderr); exit(4); }
		totR = (uint32_t)n;
		Start = xmalloc((size_t)totR * 4); Src = xmalloc((size_t)totR * 4);
		SSeq = xmalloc((size_t)totR * sizeof(*SSeq)); SLen = xmalloc((size_t)totR * 4);
		uint32_t x = 0;
		for (uint32_t i = 0; i < origR; ++i) {
			long unit = (long)Len[i] - (long)ov; if (unit <= 0) unit = 1;
			for (long j = 0; j < unit; j += shear) {
				Src[x] = i; Start[x] = (uint32_t)j; SSeq[x] = Seq[i] + j;
				uint32_t l = Len[i] - (uint32_t)j; SLen[x] = l > shear + ov ? shear + ov : l; ++x;
			}
		}
		printf("Shorn refs: %u, rebased clumps: %u\n", totR, totR / 16 + (totR % 16 != 0));
	}
	/* ---- order: length pods of LATENCY bases, lexicographic inside (as load_fasta_refs) ---- */
	Tux *T = xmalloc((size_t)totR * sizeof(*T));
	for (uint32_t i = 0; i < totR; ++i) T[i] = (Tux){SSeq[i], SLen[i], i};
	qsort(T, totR, sizeof(*T), cmp_tux_len);
	maxLenR = T[totR - 1].len;
	uint32_t prev = 0, curTol = T[0].len;
	for (uint32_t i = 1; i <= totR; ++i) if (i == totR || T[i].len > curTol + LATENCY) {
		if (i - prev > 1) qsort(T + prev, i - prev, sizeof(*T), cmp_tux_seq);
		if (i < totR) curTol = T[i].len;
		prev = i;
	}
	Refs R; memset(&R, 0, sizeof(R));
	R.totR = R.origTotR = totR; R.maxLenR = maxLenR;
	R.RefIxSrt = xmalloc(((size_t)totR + 1) * 4);
	for (uint32_t i = 0; i < totR; ++i) R.RefIxSrt[i] = T[i].ix;
	free(T);
	pack_clumps(&R, SSeq, SLen);
	printf("There are %u references and hence %u clumps\n", totR, R.numRclumps);
	/* ---- .edx (burst.c:2758-2839) ---- */
	FILE *out = fopen(edx_FN, "wb");
	if (!out) { fprintf(stderr, "ERROR: Cannot open output: %s\n", edx_FN); exit(2); }
	setvbuf(out, 0, _IOFBF, 1 << 22);
	puts("Writing database...");
	fputc(1 << 7 | REBASE << 6 | 3, out);
	/* unique sorted headers, NUL-separated; RefMap[shear] = index of its header */
	char **hs = xmalloc((size_t)origR * sizeof(*hs)); uint32_t *hmap = xmalloc((size_t)origR * 4);
	uint32_t *hix = xmalloc((size_t)origR * 4);
	for (uint32_t i = 0; i < origR; ++i) hs[i] = Head[i];
	qsort(hs, origR, sizeof(*hs), cmp_strp);
	uint32_t nheads = 0; uint64_t totRefHeadLen = 0;
	for (uint32_t i = 0; i < origR; ++i) if (!i || strcmp(hs[i], hs[nheads - 1])) { hs[nheads++] = hs[i]; totRefHeadLen += strlen(hs[i]) + 1; }
	for (uint32_t i = 0; i < origR; ++i) {                                /* binary search of each original header */
		uint32_t lo = 0, hi = nheads;
		while (lo + 1 < hi) { uint32_t mid = (lo + hi) / 2; if (strcmp(hs[mid], Head[i]) <= 0) lo = mid; else hi = mid; }
		hmap[i] = lo;
	}
	(void)hix;
	uint32_t shearHdr = (uint32_t)(dbQLen / THRES);
	fwrite(&totRefHeadLen, 8, 1, out); fwrite(&shearHdr, 4, 1, out); fwrite(&totR, 4, 1, out); fwrite(&totR, 4, 1, out);
	fwrite(&R.numRclumps, 4, 1, out); fwrite(&maxLenR, 4, 1, out);
	for (uint32_t i = 0; i < nheads; ++i) fwrite(hs[i], 1, strlen(hs[i]) + 1, out);
	fwrite(&nheads, 4, 1, out);
	for (uint32_t i = 0; i < totR; ++i) { uint32_t m = hmap[Src ? Src[i] : i]; fwrite(&m, 4, 1, out); }       /* RefMap */
	if (REBASE) fwrite(Start, 4, totR, out);                              /* RefStart */
	fwrite(R.RefIxSrt, 4, totR, out);                                     /* TmpRIX: lane slot -> sheared reference (no de-duplication here) */
	fwrite(R.ClumpLen, 4, R.numRclumps, out);
	fwrite(R.packed, 1, R.packedBytes, out);
	fclose(out);
	puts("Database written.");
	/* ---- .acx (burst.c:3304-3532) ---- */
	if (acx_FN) {
		printf("Generating accelerator '%s'\n", acx_FN);
		if (Z) fprintf(stderr, "Note: N-penalized accelerator not usable for unpenalized alignment\n");
		const uint64_t nk = 1ull << (2 * N);
		uint32_t *Lens = xcalloc(nk, 4);
		WordSet S = {xcalloc(nk >> 3, 1), xmalloc((1u << 16) * 4), 0, 1u << 16, N, N == 16 ? 0xFFFFFFFFu : (uint32_t)(nk - 1)};
		uint32_t *Bad = xmalloc(((size_t)R.numRclumps + 1) * 4), badSz = 0;
		uint8_t *isBad = xcalloc(R.numRclumps + 1, 1);
		uint64_t total = 0;
		for (uint32_t c = 0; c < R.numRclumps; ++c) {                      /* pass 1: posting-list lengths */
			uint32_t nl = MIN(VECSZ, totR - c * VECSZ);
			if (clump_words(&S, SSeq, SLen, R.RefIxSrt + (size_t)c * VECSZ, nl, skipAmbig)) { isBad[c] = 1; Bad[badSz++] = c; }
			else for (uint32_t k = 0; k < S.n; ++k) ++Lens[S.cache[k]];
			for (uint32_t k = 0; k < S.n; ++k) S.yn[S.cache[k] >> 3] = 0;
			total += isBad[c] ? 0 : S.n;
		}
		printf("Total accelerants stored: %" PRIu64 " (%u bad)\n", total, badSz);
		uint64_t *off = xmalloc((nk + 1) * 8); off[0] = 0;
		for (uint64_t i = 0; i < nk; ++i) off[i + 1] = off[i] + Lens[i];
		uint32_t *post = xmalloc((total + 1) * 4), *fill = xcalloc(nk, 4);
		for (uint32_t c = 0; c < R.numRclumps; ++c) {                      /* pass 2: postings, ascending clump id per word */
			if (isBad[c]) continue;
			uint32_t nl = MIN(VECSZ, totR - c * VECSZ);
			clump_words(&S, SSeq, SLen, R.RefIxSrt + (size_t)c * VECSZ, nl, skipAmbig);
			for (uint32_t k = 0; k < S.n; ++k) { uint32_t w = S.cache[k]; post[off[w] + fill[w]++] = c; S.yn[w >> 3] = 0; }
		}
		FILE *ax = fopen(acx_FN, "wb");
		if (!ax) { fprintf(stderr, "Cannot write accelerator '%s'\n", acx_FN); exit(1); }
		setvbuf(ax, 0, _IOFBF, 1 << 22);
		if (R.numRclumps > 16777214) { fputs("ERROR: acc error M16\n", stderr); exit(101); }
		int big = R.numRclumps > 1048574;
		fputc(1 << 7 | Z << 6 | (big ? 1 : 0), ax);
		fwrite(&badSz, 4, 1, ax);
		fwrite(Lens, 4, nk, ax);
		fprintf(stderr, " --> [Re-Accel] Writing %s format acx...\n", big ? "LARGE" : "SMALL");
		if (big) for (uint64_t i = 0; i < total; ++i) fwrite(post + i, 3, 1, ax);
		else for (uint64_t w = 0; w < nk; ++w) for (uint64_t p = off[w]; p < off[w + 1]; p += 2) {
			uint64_t bay = post[p];
			if (p + 1 < off[w + 1]) { bay |= (uint64_t)post[p + 1] << 20; fwrite(&bay, 1, 5, ax); }
			else fwrite(&bay, 1, 3, ax);
		}
		fwrite(Bad, 4, badSz, ax);
		fclose(ax);
		printf("Wrote accelerator (DB%d).\n", N);
	}
}

/* =============================================================================================
 * Accelerated search (burst.c:4018-4316).  Bunches of QBUNCH sorted queries share one candidate list
 * (burst.c:4085-4133); a bunch visits its candidates by descending k-mer count, then the always-visited
 * BadList, and inside a clump its queries in order (burst.c:4136-4168).  Here the visits of many bunches
 * form one BATCH: the (bunch, clump) visits become bg_run records in exactly that order -- so hits sorted
 * by run index are in the reference's -t 1 discovery order -- the reads of the batch are nibble-packed
 * into page-locked memory, and one bg_align_runs_into() call does the rest on the GPU.
 *   * candidate generation runs on all host threads (-t), one bunch per task, into per-thread arenas that
 *     are stitched together in bunch order;
 *   * batches are dealt to the GPUs (--gpus N, one engine context each, database replicated) in waves:
 *     while the GPUs work on wave w the host threads generate wave w+1; every GPU carries its own per-read
 *     running minima (ShrBins[].ed), merged by MIN at the end exactly like the reference's thread pods
 *     (burst.c:4497-4517) -- a read's forward and reverse-complement strands may sit in different batches;
 *   * the per-query skip (burst.c:4163-4168: count <= len - (ed+1) N) is kept exactly: a visit is cut into
 *     the maximal runs of consecutive queries that pass it.
 * ============================================================================================= */
typedef struct { bg_run *r; uint64_t n, cap; } RunVec;
static void run_push(RunVec *V, uint32_t clump, uint32_t q0, uint32_t nq) {
	if (V->n == V->cap) { V->cap = V->cap ? V->cap * 2 : (1 << 14); V->r = xrealloc(V->r, V->cap * sizeof(bg_run)); }
	V->r[V->n].clump = clump; V->r[V->n].query0 = q0; V->r[V->n++].nq = nq;
}
typedef struct {                       /* per host thread */
	uint16_t *Hash; uint32_t *Cache; Split *Cand; uint64_t *W, wcap; RunVec runs;
} GenScratch;
typedef struct { uint32_t tid; uint64_t off, n; } BunchRuns;   /* where a bunch's runs sit: thread arena, offset, count */
typedef struct {
	uint64_t qa, qb, nbunch;           /* UniBins [qa, qb), bunches */
	BunchRuns *br;                     /* per bunch */
	bg_run *runs; uint64_t nruns, runcap;
	uint8_t *codes; uint64_t *off; uint16_t *bud; uint32_t *slot; uint64_t cap_codes, cap_q;   /* page-locked */
	bg_hit *hits; uint64_t hitcap;
} Batch;

/* candidates of one bunch [z, bound) -> runs appended to S->runs; query numbers are relative to qa */
static void gen_bunch(const Queries *Q, const Acx *A, uint32_t numRclumps, GenScratch *S, uint64_t z, uint64_t bound, uint64_t qa) {
	const uint32_t N = (uint32_t)SCOUR_N;
	uint64_t wix = 0; uint64_t *W = S->W;
	uint32_t min_mmatch = UINT32_MAX, mm[16];
	for (uint64_t j = z; j < bound; ++j) {                                /* burst.c:4085-4114 */
		const UniBin *u = Q->UniBins + j; const ShrBin *sb = Q->ShrBins + u->six;
		uint32_t len = sb->len, err = sb->ed, kload = err * N + N, mmatch = kload < len ? len - kload : 0;
		uint32_t heur = DO_HEUR ? (len >> 4) + 1u : 0;
		if (mmatch < heur) mmatch = heur;
		if (mmatch < min_mmatch) min_mmatch = mmatch;
		mm[j - z] = kload < len ? len - kload : 1;                        /* the per-query skip threshold, burst.c:4163-4164 */
		const char *s = u->seq;
		uint64_t need = wix + (uint64_t)len * (j >= Q->QBins[0] ? 1 : 1024) + 16;
		if (need > S->wcap) { while (S->wcap < need) S->wcap *= 2; W = S->W = xrealloc(S->W, S->wcap * 8); }
		if (j >= Q->QBins[0]) {                                           /* unambiguous: rolling 2-bit words */
			uint32_t mask = N == 16 ? 0xFFFFFFFFu : (1u << (2 * N)) - 1, w = 0;
			for (uint32_t k = 0; k < len; ++k) {
				w = (w << 2 | (uint32_t)(s[k] - 1)) & mask;
				if (k + 1 >= N) W[wix++] = (uint64_t)w << 32 | (uint32_t)(j - z);
			}
		} else for (uint32_t k = 0; k + N <= len; ++k) {                  /* ambiguous: every variant of every window */
			if (Z) { uint32_t p = k, e = k + N; for (; p < e; ++p) if (s[p] == 5) break; if (p < e) { k = p; continue; } }
			uint64_t variants = 1;
			for (uint32_t p = k; p < k + N; ++p) variants *= AMBIG_N[(uint8_t)s[p] & 15];
			if (wix + variants + 16 > S->wcap) { while (S->wcap < wix + variants + 16) S->wcap *= 2; W = S->W = xrealloc(S->W, S->wcap * 8); }
			ambig_words(W, &wix, s + k, (uint32_t)(j - z), 0, 0);
		}
	}
	qsort(W, wix, 8, cmp_u64);                                            /* burst.c:4118 */
	/* burst.c:3238-3282: each distinct word adds its largest per-query multiplicity to every clump it lists */
	uint16_t *Hash = S->Hash; uint32_t *Cache = S->Cache, cix = 0;
	for (uint64_t i = 0; i < wix;) {
		uint32_t v = (uint32_t)(W[i] >> 32), mx = 0; uint64_t e = i;
		while (e < wix && (uint32_t)(W[e] >> 32) == v) { uint64_t r = e; while (r < wix && W[r] == W[e]) ++r; if (r - e > mx) mx = (uint32_t)(r - e); e = r; }
		const uint8_t *p = A->post + A->off[v], *end = A->post + A->off[(uint64_t)v + 1];
#define BUMP(px) do { uint32_t px_ = (px); if (px_ < numRclumps) { if (!Hash[px_]) Cache[cix++] = px_; uint32_t nv_ = mx + Hash[px_]; Hash[px_] = (uint16_t)(nv_ > 65535 ? 65535 : nv_); } } while (0)
		if (A->big) for (; p < end; p += 3) { uint32_t px; memcpy(&px, p, 4); BUMP(px & 0xFFFFFF); }
		else for (; p < end; p += 5) {
			uint64_t PX; memcpy(&PX, p, 8);
			BUMP((uint32_t)(PX & 0xFFFFF));
			if (p + 3 >= end) break;
			BUMP((uint32_t)((PX >> 20) & 0xFFFFF));
		}
		i = e;
	}
	Split *Cand = S->Cand; uint32_t nref = 0;
	for (uint32_t i = 0; i < cix; ++i) { uint16_t *h = Hash + Cache[i]; if (*h > min_mmatch) Cand[nref++] = (Split){Cache[i], *h}; *h = 0; }
	if (nref > 24) qsort(Cand, nref, sizeof(*Cand), cmp_refcount);        /* burst.c:4038-4046 */
	else for (uint32_t i = 1, j; i < nref; ++i) {
		Split key = Cand[i];
		for (j = i; j && Cand[j - 1].i < key.i; --j);
		memmove(Cand + j + 1, Cand + j, sizeof(*Cand) * (i - j)); Cand[j] = key;
	}
	const uint32_t nb = (uint32_t)(bound - z), q0 = (uint32_t)(z - qa);
	for (uint32_t i = 0; i < nref; ++i) {                                 /* burst.c:4137-4168 */
		uint32_t a = 0;
		while (a < nb) {
			while (a < nb && !(Cand[i].i > mm[a])) ++a;
			uint32_t b = a;
			while (b < nb && Cand[i].i > mm[b]) ++b;
			if (b > a) run_push(&S->runs, Cand[i].v, q0 + a, b - a);
			a = b;
		}
	}
	if (!Q->skipAmbig) for (uint32_t i = 0; i < A->nbad; ++i) if (A->bad[i] < numRclumps) run_push(&S->runs, A->bad[i], q0, nb);
}

static void batch_room(Batch *B, uint64_t nq, uint64_t ncodes, uint64_t nruns) {
	if (nq + 1 > B->cap_q) {
		bg_host_free(B->off); bg_host_free(B->bud); bg_host_free(B->slot);
		B->cap_q = nq + nq / 8 + 64;
		B->off = bg_host_alloc(B->cap_q * 8); B->bud = bg_host_alloc(B->cap_q * 2); B->slot = bg_host_alloc(B->cap_q * 4);
	}
	if (ncodes / 2 + 64 > B->cap_codes) { bg_host_free(B->codes); B->cap_codes = ncodes / 2 + ncodes / 16 + 256; B->codes = bg_host_alloc(B->cap_codes); }
	if (nruns + 1 > B->runcap) { bg_host_free(B->runs); B->runcap = nruns + nruns / 8 + 64; B->runs = bg_host_alloc(B->runcap * sizeof(bg_run)); }
	if (!B->off || !B->bud || !B->slot || !B->codes || !B->runs) { fputs("OOM: page-locked batch buffers\n", stderr); exit(3); }
}

/* queries of the batch -> nibble-packed codes (two bases per byte, even base low: BG_Q_PACKED4), offsets in bases */
static void batch_pack_queries(const Queries *Q, Batch *B) {
	uint64_t nq = B->qb - B->qa, tot = 0;
	for (uint64_t j = 0; j < nq; ++j) { B->off[j] = tot; tot += Q->ShrBins[Q->UniBins[B->qa + j].six].len; }
	B->off[nq] = tot;
	memset(B->codes, 0, tot / 2 + 16);
	#pragma omp parallel for schedule(static, 4096) num_threads(THREADS)
	for (uint64_t j = 0; j < nq; ++j) {
		const UniBin *u = Q->UniBins + B->qa + j; const ShrBin *sb = Q->ShrBins + u->six;
		B->bud[j] = sb->ed; B->slot[j] = (uint32_t)u->six;
		uint64_t o = B->off[j]; const char *s = u->seq; uint32_t len = sb->len, k = 0;
		/* the first and last nibble of a query may share a byte with its neighbours (handled by another thread): atomic OR there */
		if ((o & 1) && len) { __atomic_fetch_or(&B->codes[o >> 1], (uint8_t)((s[0] & 15) << 4), __ATOMIC_RELAXED); k = 1; }
		for (; k + 1 < len; k += 2) B->codes[(o + k) >> 1] = (uint8_t)((s[k] & 15) | ((s[k + 1] & 15) << 4));
		if (k < len) __atomic_fetch_or(&B->codes[(o + k) >> 1], (uint8_t)(s[k] & 15), __ATOMIC_RELAXED);
	}
}

static void accel_search(bg_ctx **ctxs, int ngpu, Queries *Q, Refs *R, Acx *A, PodList *Pods, int mode, int threads) {
	uint64_t nAcc = Q->QBins[1], newUniqQ = Q->newUniqQ;
	if (!nAcc) return;
	uint64_t QBUNCH = newUniqQ / ((uint64_t)threads * 128);
	if (QBUNCH > 16) QBUNCH = 16;
	if (!QBUNCH) QBUNCH = 1;
	printf("Setting QBUNCH to %" PRIu64 "\nUsing ACCELERATOR to align %" PRIu64 " unique queries...\n", QBUNCH, nAcc);
	const uint32_t numRclumps = R->numRclumps;
	const uint64_t nbunch = (nAcc + QBUNCH - 1) / QBUNCH;
	/* bunches per batch: enough batches to keep every GPU busy and the generation a wave ahead, at most ~1 M queries each */
	uint64_t bpb = (nbunch + (uint64_t)ngpu * 4 - 1) / ((uint64_t)ngpu * 4);
	if (bpb < 2048) bpb = 2048;
	if (bpb > 65536) bpb = 65536;
	const uint64_t nbatch = (nbunch + bpb - 1) / bpb;
	int nthr = threads < 1 ? 1 : threads;
#ifdef _OPENMP
	if (nthr < ngpu) nthr = ngpu;
#else
	nthr = 1;
#endif
	GenScratch *GS = xcalloc((size_t)nthr, sizeof(*GS));
	for (int t = 0; t < nthr; ++t) {
		GS[t].Hash = xcalloc(numRclumps, sizeof(uint16_t)); GS[t].Cache = xmalloc(((uint64_t)numRclumps + 1) * 4);
		GS[t].Cand = xmalloc(((uint64_t)numRclumps + 1) * sizeof(Split)); GS[t].wcap = 1 << 16; GS[t].W = xmalloc(GS[t].wcap * 8);
	}
	uint16_t **best = xmalloc((size_t)ngpu * sizeof(*best));
	for (int g = 0; g < ngpu; ++g) { best[g] = xmalloc(Q->numUniqQ * sizeof(uint16_t)); for (uint64_t i = 0; i < Q->numUniqQ; ++i) best[g][i] = 0xFFFF; }
	PodList *SPods = xcalloc(newUniqQ, sizeof(*SPods));                 /* per strand, folded below (4299-4312) */
	Batch *BT = xcalloc((size_t)ngpu * 2, sizeof(*BT));                  /* two waves of ngpu batches */
	int failed = 0, fail_rc = 0;
	double t_gen = 0, t_gpu = 0;

	const uint64_t nwaves = (nbatch + (uint64_t)ngpu - 1) / (uint64_t)ngpu;
	for (uint64_t wave = 0; wave <= nwaves; ++wave) {
		/* this pass: the GPUs align wave-1 (if any) while the whole team generates the candidates of `wave` (if any) */
		const uint64_t wb0 = MIN(nbunch, wave * (uint64_t)ngpu * bpb), wb1 = MIN(nbunch, (wave + 1) * (uint64_t)ngpu * bpb);   /* bunches of this wave */
		Batch *Gen = BT + (wave & 1) * (uint64_t)ngpu, *Run = BT + ((wave + 1) & 1) * (uint64_t)ngpu;
		for (int g = 0; g < ngpu; ++g) {
			Batch *B = Gen + g; uint64_t b0 = MIN(nbunch, wb0 + (uint64_t)g * bpb), b1 = MIN(nbunch, b0 + bpb);
			B->nbunch = b1 - b0; B->qa = b0 * QBUNCH; B->qb = MIN(nAcc, b1 * QBUNCH);
			if (B->nbunch) B->br = xrealloc(B->br, B->nbunch * sizeof(BunchRuns));
		}
		for (int t = 0; t < nthr; ++t) GS[t].runs.n = 0;
		double t0 = now();
		#pragma omp parallel num_threads(nthr)
		{
			if (wave > 0) {
				#pragma omp for schedule(static, 1) nowait
				for (int g = 0; g < ngpu; ++g) {
					Batch *B = Run + g;
					if (!B->nbunch || failed) continue;
					uint64_t nq = B->qb - B->qa;
					bg_queries bq = {B->codes, B->off, B->bud, B->slot, (uint32_t)nq, (uint32_t)Q->numUniqQ, BG_Q_PACKED4};
					uint64_t nh = 0; int rc;
					if (!B->hits) { B->hitcap = nq * 2 + 1024; B->hits = bg_host_alloc(B->hitcap * sizeof(bg_hit)); }
					rc = B->nruns ? bg_align_runs_into(ctxs[g], &bq, B->runs, B->nruns, mode, best[g], B->hits, B->hitcap, &nh) : BG_OK;
					if (rc == BG_EOVERFLOW && nh > B->hitcap) {            /* more hits than room: grow and redo the batch */
						bg_host_free(B->hits); B->hitcap = nh + nh / 8; B->hits = bg_host_alloc(B->hitcap * sizeof(bg_hit));
						rc = bg_align_runs_into(ctxs[g], &bq, B->runs, B->nruns, mode, best[g], B->hits, B->hitcap, &nh);
					}
					if (rc) {
						#pragma omp critical
						{ failed = 1; fail_rc = rc; fprintf(stderr, "ERROR: GPU engine failed in bg_align_runs_into: %s\n", bg_last_error()); }
						continue;
					}
					for (uint64_t h = 0; h < nh; ++h) {
						const bg_run *r = B->runs + (B->hits[h].task >> 4);
						const UniBin *u = Q->UniBins + B->qa + r->query0 + (B->hits[h].task & 15); const ShrBin *sb = Q->ShrBins + u->six;
						uint32_t refIx = r->clump * VECSZ + B->hits[h].lane;
						if (refIx >= R->totR) continue;                     /* burst.c:4229 */
						Pod p = {identity(B->hits[h].ed, sb->len, B->hits[h].gap_q), refIx, B->hits[h].final_pos, B->hits[h].gap_r, B->hits[h].gap_q, B->hits[h].ed, u->rc};
						pod_push(SPods + u->six + (u->rc ? Q->numUniqQ : 0), p);
					}
				}
			}
			/* no barrier up to here: the threads not driving a GPU start on the bunches at once, the others join when their call returns */
			#pragma omp for schedule(dynamic, 16)
			for (uint64_t b = wb0; b < wb1; ++b) {
				int tid = 0;
#ifdef _OPENMP
				tid = omp_get_thread_num();
#endif
				GenScratch *S = GS + tid; Batch *B = Gen + (b - wb0) / bpb; uint64_t before = S->runs.n;
				gen_bunch(Q, A, numRclumps, S, b * QBUNCH, MIN(nAcc, (b + 1) * QBUNCH), B->qa);
				B->br[(b - wb0) % bpb] = (BunchRuns){(uint32_t)tid, before, S->runs.n - before};
			}
		}
		if (failed) exit(fail_rc == BG_ENOMEM ? 3 : 4);
		double t1 = now();
		/* stitch the wave's runs together in bunch order and pack its reads (page-locked buffers) */
		for (int g = 0; g < ngpu; ++g) {
			Batch *B = Gen + g;
			if (!B->nbunch) continue;
			uint64_t tot = 0, nc = 0, o = 0;
			for (uint64_t b = 0; b < B->nbunch; ++b) tot += B->br[b].n;
			for (uint64_t j = B->qa; j < B->qb; ++j) nc += Q->ShrBins[Q->UniBins[j].six].len;
			batch_room(B, B->qb - B->qa, nc, tot);
			for (uint64_t b = 0; b < B->nbunch; ++b) { memcpy(B->runs + o, GS[B->br[b].tid].runs.r + B->br[b].off, B->br[b].n * sizeof(bg_run)); o += B->br[b].n; }
			B->nruns = tot;
			batch_pack_queries(Q, B);
		}
		t_gen += now() - t1; t_gpu += t1 - t0;
		if (!QUIET) printf("\rSearch Progress: [%3.2f%%]", 100.0 * (double)MIN(wave, nwaves) / (double)(nwaves ? nwaves : 1));
	}
	printf(" --> [Accel] %" PRIu64 " batches on %d GPU(s): align + candidate generation %.3f s, stitch + pack %.3f s\n", nbatch, ngpu, t_gpu, t_gen);
	if (!QUIET) printf("\rSearch Progress: [100.00%%]\n");
	/* per-read minima over all GPUs (burst.c:4497-4517), then fold strands: forward list, then reverse-complement list (4299-4312) */
	for (int g = 1; g < ngpu; ++g) for (uint64_t i = 0; i < Q->numUniqQ; ++i) if (best[g][i] < best[0][i]) best[0][i] = best[g][i];
	for (uint64_t i = 0; i < Q->numUniqQ; ++i) {
		PodList *F = SPods + i, *Rc = Q->rc ? SPods + Q->numUniqQ + i : NULL;
		/* Pods[] is kept in discovery order and read backwards; "fwd list then rc list" read backwards is
		 * rc pods (discovery order) followed by fwd pods (discovery order) */
		if (Rc) for (uint32_t k = 0; k < Rc->n; ++k) if (mode != BG_MODE_MIN || Rc->p[k].mismatches <= best[0][i]) pod_push(Pods + i, Rc->p[k]);
		for (uint32_t k = 0; k < F->n; ++k) if (mode != BG_MODE_MIN || F->p[k].mismatches <= best[0][i]) pod_push(Pods + i, F->p[k]);
		free(F->p); if (Rc) free(Rc->p);
	}
	for (int t = 0; t < nthr; ++t) { free(GS[t].Hash); free(GS[t].Cache); free(GS[t].Cand); free(GS[t].W); free(GS[t].runs.r); }
	for (int b = 0; b < ngpu * 2; ++b) { Batch *B = BT + b; free(B->br); bg_host_free(B->runs); bg_host_free(B->codes); bg_host_free(B->off); bg_host_free(B->bud); bg_host_free(B->slot); bg_host_free(B->hits); }
	for (int g = 0; g < ngpu; ++g) free(best[g]);
	free(best); free(GS); free(BT); free(SPods);
}

/* =============================================================================================
 * --device-candidates: the accelerator lives on the GPU (bg_load_acx) and a batch is just its reads at 2 bits
 * per base plus one word per strand; the candidates of every bunch are counted, ranked and cut into runs by the device
 * (bg_search_bunches_into), which returns hits that name strand and clump.  Same bunches, same candidate rule; among clumps of EQUAL
 * count the device keeps first-touch order, which is the reference's order for lists of <= 24 clumps (its insertion sort) and for
 * longer lists whatever the C library's qsort leaves (burst.c:4038-4046) -- so BEST/CAPITALIST may pick a different one of several
 * equally good references there; ALLPATHS and FORAGE rows are the same set.  Queries with ambiguous bases need every variant of
 * every window (burst.c:3232-3236): when the accelerated bin holds any, the host lists are used instead.
 * ============================================================================================= */
typedef struct {
	uint64_t qa, qb; uint32_t nreads;
	uint8_t *reads; uint16_t *len, *bud, *best; uint32_t *strand; uint64_t *roff; uint64_t *six;      /* page-locked: reads/len/bud/strand/best */
	const char **src; uint8_t *srcrc; uint64_t cap_r, cap_q, nh;
	bg_xhit *hits; uint64_t hitcap;
} DevBatch;
static void accel_search_device(bg_ctx **ctxs, int ngpu, Queries *Q, Refs *R, Acx *A, PodList *Pods, int mode, int threads) {
	uint64_t nAcc = Q->QBins[1], newUniqQ = Q->newUniqQ;
	if (!nAcc) return;
	uint64_t QBUNCH = newUniqQ / ((uint64_t)threads * 128);
	if (QBUNCH > 16) QBUNCH = 16;
	if (!QBUNCH) QBUNCH = 1;
	printf("Setting QBUNCH to %" PRIu64 "\nUsing ACCELERATOR (on the GPU) to align %" PRIu64 " unique queries...\n", QBUNCH, nAcc);
	double t0 = now();
	for (int g = 0; g < ngpu; ++g) { int rc = bg_load_acx(ctxs[g], A->lens, A->post, A->post_bytes, SCOUR_N, A->big, A->bad, A->nbad); if (rc) die_gpu("bg_load_acx", rc); }
	free(A->lens); A->lens = NULL;
	double t_load = now() - t0;
	const uint64_t nbunch = (nAcc + QBUNCH - 1) / QBUNCH;
	uint64_t bpb = (nbunch + (uint64_t)ngpu * 4 - 1) / ((uint64_t)ngpu * 4);
	if (bpb < 2048) bpb = 2048;
	if (bpb > 65536) bpb = 65536;
	const uint64_t nbatch = (nbunch + bpb - 1) / bpb;
	uint16_t *best = xmalloc(Q->numUniqQ * sizeof(*best));
	for (uint64_t i = 0; i < Q->numUniqQ; ++i) best[i] = 0xFFFF;
	uint32_t *lid = xmalloc(Q->numUniqQ * 4), *stamp = xcalloc(Q->numUniqQ, 4);
	PodList *SPods = xcalloc(newUniqQ, sizeof(*SPods));
	DevBatch *BT = xcalloc((size_t)ngpu, sizeof(*BT));
	double t_pack = 0, t_gpu = 0;
	for (uint64_t w0 = 0; w0 < nbatch; w0 += (uint64_t)ngpu) {
		double ta = now();
		int nb_wave = (int)MIN((uint64_t)ngpu, nbatch - w0);
		for (int g = 0; g < nb_wave; ++g) {                               /* the batch's distinct reads, in order of first appearance */
			DevBatch *B = BT + g; uint64_t b0 = (w0 + (uint64_t)g) * bpb, b1 = MIN(nbunch, b0 + bpb);
			B->qa = b0 * QBUNCH; B->qb = MIN(nAcc, b1 * QBUNCH);
			uint64_t nq = B->qb - B->qa;
			if (nq + 1 > B->cap_q) {
				bg_host_free(B->strand); bg_host_free(B->len); bg_host_free(B->bud); bg_host_free(B->best); free(B->roff); free(B->six); free(B->src); free(B->srcrc);
				B->cap_q = nq + nq / 8 + 64;
				B->strand = bg_host_alloc(B->cap_q * 4); B->len = bg_host_alloc(B->cap_q * 2); B->bud = bg_host_alloc(B->cap_q * 2); B->best = bg_host_alloc(B->cap_q * 2);
				B->roff = xmalloc((B->cap_q + 1) * 8); B->six = xmalloc(B->cap_q * 8); B->src = xmalloc(B->cap_q * sizeof(*B->src)); B->srcrc = xmalloc(B->cap_q);
				if (!B->strand || !B->len || !B->bud || !B->best) { fputs("OOM: page-locked batch buffers\n", stderr); exit(3); }
			}
			const uint32_t epoch = (uint32_t)(w0 + (uint64_t)g) + 1; uint32_t nr = 0; uint64_t tot = 0;
			for (uint64_t j = 0; j < nq; ++j) {
				const UniBin *u = Q->UniBins + B->qa + j; const ShrBin *sb = Q->ShrBins + u->six;
				if (stamp[u->six] != epoch) {
					stamp[u->six] = epoch; lid[u->six] = nr;
					B->len[nr] = (uint16_t)sb->len; B->bud[nr] = sb->ed; B->best[nr] = best[u->six]; B->six[nr] = u->six; B->src[nr] = u->seq; B->srcrc[nr] = u->rc;
					B->roff[nr] = tot; tot += sb->len; ++nr;
				}
				B->strand[j] = lid[u->six] | (u->rc ? 0x80000000u : 0u);
			}
			B->roff[nr] = tot; B->nreads = nr;
			if (tot / 4 + 64 > B->cap_r) { bg_host_free(B->reads); B->cap_r = tot / 4 + tot / 32 + 256; B->reads = bg_host_alloc(B->cap_r); if (!B->reads) { fputs("OOM: page-locked batch buffers\n", stderr); exit(3); } }
			memset(B->reads, 0, tot / 4 + 16);
			#pragma omp parallel for schedule(static, 4096) num_threads(threads < 1 ? 1 : threads)
			for (uint32_t r = 0; r < nr; ++r) {                              /* forward read at 2 bits per base; a read first met through its reverse complement is turned back */
				const char *sq = B->src[r]; const uint32_t len = B->len[r]; const int rcs = B->srcrc[r]; uint64_t o = B->roff[r];
				for (uint32_t k = 0; k < len; ++k, ++o) {
					const uint8_t code = rcs ? (uint8_t)(5 - sq[len - 1 - k]) : (uint8_t)sq[k];
					const uint8_t bits = (uint8_t)(((code - 1) & 3) << (2 * (o & 3)));
					if (k < 4 || k + 4 >= len) __atomic_fetch_or(&B->reads[o >> 2], bits, __ATOMIC_RELAXED); else B->reads[o >> 2] |= bits;
				}
			}
		}
		double tb = now(); t_pack += tb - ta;
		int failed = 0;
		#pragma omp parallel for schedule(static, 1) num_threads(nb_wave)
		for (int g = 0; g < nb_wave; ++g) {
			DevBatch *B = BT + g; uint64_t nq = B->qb - B->qa, nh = 0;
			bg_reads br = {B->reads, B->len, B->bud, B->strand, B->nreads, (uint32_t)nq, BG_R_PACKED2};
			if (!B->hits) { B->hitcap = nq * 2 + 1024; B->hits = bg_host_alloc(B->hitcap * sizeof(bg_xhit)); }
			int rc = bg_search_bunches_into(ctxs[g], &br, (uint32_t)QBUNCH, DO_HEUR, Q->skipAmbig, mode, B->best, B->hits, B->hitcap, &nh);
			if (rc == BG_EOVERFLOW && nh > B->hitcap) {
				bg_host_free(B->hits); B->hitcap = nh + nh / 8; B->hits = bg_host_alloc(B->hitcap * sizeof(bg_xhit));
				for (uint32_t r = 0; r < B->nreads; ++r) B->best[r] = best[B->six[r]];
				rc = bg_search_bunches_into(ctxs[g], &br, (uint32_t)QBUNCH, DO_HEUR, Q->skipAmbig, mode, B->best, B->hits, B->hitcap, &nh);
			}
			if (rc) {
				#pragma omp critical
				{ failed = rc; fprintf(stderr, "ERROR: GPU engine failed in bg_search_bunches_into: %s\n", bg_last_error()); }
				continue;
			}
			B->nh = nh;
		}
		if (failed) exit(failed == BG_ENOMEM ? 3 : 4);
		for (int g = 0; g < nb_wave; ++g) {                               /* fold in batch order: per strand the hits arrive in visiting order */
			DevBatch *B = BT + g;
			for (uint32_t r = 0; r < B->nreads; ++r) if (B->best[r] < best[B->six[r]]) best[B->six[r]] = B->best[r];
			for (uint64_t h = 0; h < B->nh; ++h) {
				const bg_xhit *x = B->hits + h;
				const UniBin *u = Q->UniBins + B->qa + x->query; const ShrBin *sb = Q->ShrBins + u->six;
				uint32_t refIx = x->clump * VECSZ + x->lane;
				if (refIx >= R->totR) continue;                                 /* burst.c:4229 */
				Pod p = {identity(x->ed, sb->len, x->gap_q), refIx, x->final_pos, x->gap_r, x->gap_q, x->ed, u->rc};
				pod_push(SPods + u->six + (u->rc ? Q->numUniqQ : 0), p);
			}
		}
		t_gpu += now() - tb;
		if (!QUIET) printf("\rSearch Progress: [%3.2f%%]", 100.0 * (double)MIN(w0 + (uint64_t)ngpu, nbatch) / (double)nbatch);
	}
	if (!QUIET) printf("\rSearch Progress: [100.00%%]\n");
	printf(" --> [Accel] %" PRIu64 " batches on %d GPU(s), candidates on the device: accelerator upload %.3f s, read packing %.3f s, search + fold %.3f s\n", nbatch, ngpu, t_load, t_pack, t_gpu);
	for (uint64_t i = 0; i < Q->numUniqQ; ++i) {
		PodList *F = SPods + i, *Rc = Q->rc ? SPods + Q->numUniqQ + i : NULL;
		if (Rc) for (uint32_t k = 0; k < Rc->n; ++k) if (mode != BG_MODE_MIN || Rc->p[k].mismatches <= best[i]) pod_push(Pods + i, Rc->p[k]);
		for (uint32_t k = 0; k < F->n; ++k) if (mode != BG_MODE_MIN || F->p[k].mismatches <= best[i]) pod_push(Pods + i, F->p[k]);
		free(F->p); if (Rc) free(Rc->p);
	}
	for (int g = 0; g < ngpu; ++g) { DevBatch *B = BT + g; bg_host_free(B->reads); bg_host_free(B->strand); bg_host_free(B->len); bg_host_free(B->bud); bg_host_free(B->best); bg_host_free(B->hits); free(B->roff); free(B->six); free(B->src); free(B->srcrc); }
	free(BT); free(best); free(lid); free(stamp); free(SPods);
}

#ifdef BURST_NCCL
/* =============================================================================================
 * Reference sharding (--shard-refs): the .edx does not fit one GPU, so GPU g holds the clump range
 * [lo_g, hi_g); every GPU sees every batch and skips the runs of foreign clumps.  Per batch: filter + extend on every
 * GPU, then ONE collective -- ncclAllReduce(MIN) over the per-read minima (uint32 x numUniqQ, on each engine's own
 * stream, no host synchronisation in between) -- then the selection against the combined minima: exactly the rule by
 * which the reference merges its thread pods (burst.c:4490-4519).  FORAGE keeps every lane within budget and needs no
 * collective.  One host thread per GPU; the all-reduce is NCCL over NVLink.
 * ============================================================================================= */
static int cmp_hit(const void *a, const void *b) {
	const bg_hit *A = a, *B = b;
	if (A->task != B->task) return A->task < B->task ? -1 : 1;
	return (int)A->lane - (int)B->lane;
}
static void accel_search_sharded(bg_ctx **ctxs, int ngpu, Queries *Q, Refs *R, Acx *A, PodList *Pods, int mode, int threads) {
	uint64_t nAcc = Q->QBins[1], newUniqQ = Q->newUniqQ;
	if (!nAcc) return;
	uint64_t QBUNCH = newUniqQ / ((uint64_t)threads * 128);
	if (QBUNCH > 16) QBUNCH = 16;
	if (!QBUNCH) QBUNCH = 1;
	printf("Setting QBUNCH to %" PRIu64 "\nUsing ACCELERATOR to align %" PRIu64 " unique queries on %d reference shards...\n", QBUNCH, nAcc, ngpu);
	ncclComm_t comms[64]; int devs[64];
	for (int g = 0; g < ngpu; ++g) devs[g] = GPU_DEVICE + g;
	if (ncclCommInitAll(comms, ngpu, devs) != ncclSuccess) { fputs("ERROR: ncclCommInitAll failed\n", stderr); exit(4); }
	const uint32_t numRclumps = R->numRclumps;
	const uint64_t nbunch = (nAcc + QBUNCH - 1) / QBUNCH;
	uint64_t bpb = 65536;
	const uint64_t nbatch = (nbunch + bpb - 1) / bpb;
	int nthr = threads < ngpu ? ngpu : threads;
	GenScratch *GS = xcalloc((size_t)nthr, sizeof(*GS));
	for (int t = 0; t < nthr; ++t) {
		GS[t].Hash = xcalloc(numRclumps, sizeof(uint16_t)); GS[t].Cache = xmalloc(((uint64_t)numRclumps + 1) * 4);
		GS[t].Cand = xmalloc(((uint64_t)numRclumps + 1) * sizeof(Split)); GS[t].wcap = 1 << 16; GS[t].W = xmalloc(GS[t].wcap * 8);
	}
	uint16_t *best = xmalloc(Q->numUniqQ * sizeof(*best));
	for (uint64_t i = 0; i < Q->numUniqQ; ++i) best[i] = 0xFFFF;
	PodList *SPods = xcalloc(newUniqQ, sizeof(*SPods));
	Batch B; memset(&B, 0, sizeof(B));
	bg_hit **H = xcalloc((size_t)ngpu, sizeof(*H)); uint64_t *NH = xcalloc((size_t)ngpu, sizeof(*NH));
	int failed = 0; double t_all = 0, t_gen = 0;
	for (uint64_t bi = 0; bi < nbatch; ++bi) {
		uint64_t b0 = bi * bpb, b1 = MIN(nbunch, b0 + bpb);
		B.nbunch = b1 - b0; B.qa = b0 * QBUNCH; B.qb = MIN(nAcc, b1 * QBUNCH);
		B.br = xrealloc(B.br, B.nbunch * sizeof(BunchRuns));
		for (int t = 0; t < nthr; ++t) GS[t].runs.n = 0;
		double t0 = now();
		#pragma omp parallel for schedule(dynamic, 16) num_threads(nthr)
		for (uint64_t b = b0; b < b1; ++b) {
			int tid = 0;
#ifdef _OPENMP
			tid = omp_get_thread_num();
#endif
			GenScratch *S = GS + tid; uint64_t before = S->runs.n;
			gen_bunch(Q, A, numRclumps, S, b * QBUNCH, MIN(nAcc, (b + 1) * QBUNCH), B.qa);
			B.br[b - b0] = (BunchRuns){(uint32_t)tid, before, S->runs.n - before};
		}
		uint64_t tot = 0, nc = 0, o = 0;
		for (uint64_t b = 0; b < B.nbunch; ++b) tot += B.br[b].n;
		for (uint64_t j = B.qa; j < B.qb; ++j) nc += Q->ShrBins[Q->UniBins[j].six].len;
		batch_room(&B, B.qb - B.qa, nc, tot);
		for (uint64_t b = 0; b < B.nbunch; ++b) { memcpy(B.runs + o, GS[B.br[b].tid].runs.r + B.br[b].off, B.br[b].n * sizeof(bg_run)); o += B.br[b].n; }
		B.nruns = tot;
		batch_pack_queries(Q, &B);
		double t1 = now(); t_gen += t1 - t0;
		uint64_t nq = B.qb - B.qa;
		bg_queries bq = {B.codes, B.off, B.bud, B.slot, (uint32_t)nq, (uint32_t)Q->numUniqQ, BG_Q_PACKED4};
		uint16_t *best_out = xmalloc(Q->numUniqQ * sizeof(*best_out));
		#pragma omp parallel num_threads(ngpu)
		{
			int g = 0;
#ifdef _OPENMP
			g = omp_get_thread_num();
#endif
			int rc = B.nruns ? bg_batch_upload_runs(ctxs[g], &bq, B.runs, B.nruns) : BG_OK;
			if (!rc && B.nruns) rc = bg_batch_run_extend(ctxs[g], mode, best);
			if (rc) {
				#pragma omp critical
				{ failed = rc; fprintf(stderr, "ERROR: GPU engine failed on shard %d: %s\n", g, bg_last_error()); }
			}
			#pragma omp barrier
			if (!failed && B.nruns) {
				if (mode == BG_MODE_MIN && ncclAllReduce(bg_batch_best_device(ctxs[g]), bg_batch_best_device(ctxs[g]), Q->numUniqQ, ncclUint32, ncclMin, comms[g], (cudaStream_t)bg_stream(ctxs[g])) != ncclSuccess) {
					#pragma omp critical
					{ failed = 4; fprintf(stderr, "ERROR: ncclAllReduce failed on shard %d\n", g); }
				}
				uint64_t n = 0;
				if (!failed) rc = bg_batch_run_select(ctxs[g], mode);
				if (!failed && !rc) rc = bg_batch_count(ctxs[g], &n);
				if (!failed && !rc) { H[g] = xrealloc(H[g], (n + 1) * sizeof(bg_hit)); rc = bg_batch_download(ctxs[g], H[g], n, g == 0 ? best_out : NULL); NH[g] = n; }
				if (rc) {
					#pragma omp critical
					{ failed = rc; fprintf(stderr, "ERROR: GPU engine failed on shard %d: %s\n", g, bg_last_error()); }
				}
			} else NH[g] = 0;
		}
		if (failed) exit(failed == BG_ENOMEM ? 3 : 4);
		if (B.nruns) {
			if (mode == BG_MODE_MIN) memcpy(best, best_out, Q->numUniqQ * sizeof(*best));
			else for (uint64_t i = 0; i < Q->numUniqQ; ++i) best[i] = MIN(best[i], best_out[i]);
			uint64_t nh = 0;
			for (int g = 0; g < ngpu; ++g) nh += NH[g];
			bg_hit *all = xmalloc((nh + 1) * sizeof(bg_hit)); nh = 0;
			for (int g = 0; g < ngpu; ++g) { memcpy(all + nh, H[g], NH[g] * sizeof(bg_hit)); nh += NH[g]; }
			qsort(all, nh, sizeof(bg_hit), cmp_hit);                       /* a run lives on one shard only: (task, lane) order = discovery order */
			for (uint64_t h = 0; h < nh; ++h) {
				const bg_run *r = B.runs + (all[h].task >> 4);
				const UniBin *u = Q->UniBins + B.qa + r->query0 + (all[h].task & 15); const ShrBin *sb = Q->ShrBins + u->six;
				uint32_t refIx = r->clump * VECSZ + all[h].lane;
				if (refIx >= R->totR) continue;
				Pod p = {identity(all[h].ed, sb->len, all[h].gap_q), refIx, all[h].final_pos, all[h].gap_r, all[h].gap_q, all[h].ed, u->rc};
				pod_push(SPods + u->six + (u->rc ? Q->numUniqQ : 0), p);
			}
			free(all);
		}
		free(best_out);
		t_all += now() - t1;
		if (!QUIET) printf("\rSearch Progress: [%3.2f%%]", 100.0 * (double)(bi + 1) / (double)nbatch);
	}
	if (!QUIET) printf("\rSearch Progress: [100.00%%]\n");
	printf(" --> [Accel] %" PRIu64 " batches on %d reference shards: candidate generation %.3f s, align + all-reduce(MIN) + select %.3f s\n", nbatch, ngpu, t_gen, t_all);
	for (uint64_t i = 0; i < Q->numUniqQ; ++i) {
		PodList *F = SPods + i, *Rc = Q->rc ? SPods + Q->numUniqQ + i : NULL;
		if (Rc) for (uint32_t k = 0; k < Rc->n; ++k) if (mode != BG_MODE_MIN || Rc->p[k].mismatches <= best[i]) pod_push(Pods + i, Rc->p[k]);
		for (uint32_t k = 0; k < F->n; ++k) if (mode != BG_MODE_MIN || F->p[k].mismatches <= best[i]) pod_push(Pods + i, F->p[k]);
		free(F->p); if (Rc) free(Rc->p);
	}
	for (int t = 0; t < nthr; ++t) { free(GS[t].Hash); free(GS[t].Cache); free(GS[t].Cand); free(GS[t].W); free(GS[t].runs.r); }
	for (int g = 0; g < ngpu; ++g) { free(H[g]); ncclCommDestroy(comms[g]); }
	free(B.br); bg_host_free(B.runs); bg_host_free(B.codes); bg_host_free(B.off); bg_host_free(B.bud); bg_host_free(B.slot);
	free(H); free(NH); free(best); free(GS); free(SPods);
}
#endif

/* =============================================================================================
 * Search: all-vs-all (no accelerator, and the accelerator's left-over bin), burst.c:4318-4520.
 * Discovery order per query slot there is (clump, query, lane) ascending; the engine returns
 * hits sorted by (task, lane) with task = clump * nq + query, which is the same order.
 * ============================================================================================= */
static void search_all_vs_all(bg_ctx *ctx, Queries *Q, Refs *R, uint64_t firstQ, PodList *Pods, int mode) {
	uint64_t nqAll = Q->newUniqQ - firstQ;
	if (!nqAll) return;
	printf("Searching best paths through %" PRIu64 " unique queries...\n", nqAll);
	uint16_t *best = xmalloc(Q->numUniqQ * sizeof(*best));
	for (uint64_t i = 0; i < Q->numUniqQ; ++i) best[i] = 0xFFFF;
	/* batch over queries so that one batch stays below 2^32 tasks and a few hundred MB of codes */
	uint64_t maxq = ((1ull << 32) - 2) / (R->numRclumps ? R->numRclumps : 1);
	if (maxq > (1u << 22)) maxq = 1u << 22;
	if (!maxq) { fputs("ERROR: database has too many clumps for all-vs-all search\n", stderr); exit(4); }
	for (uint64_t z = firstQ; z < Q->newUniqQ; z += maxq) {
		uint64_t bound = MIN(z + maxq, Q->newUniqQ), nq = bound - z, tot = 0;
		uint64_t *off = xmalloc((nq + 1) * sizeof(*off));
		uint16_t *bud = xmalloc(nq * sizeof(*bud)); uint32_t *slot = xmalloc(nq * sizeof(*slot));
		for (uint64_t j = 0; j < nq; ++j) { off[j] = tot; tot += Q->ShrBins[Q->UniBins[z + j].six].len; }
		off[nq] = tot;
		uint8_t *codes = xmalloc(tot + 16);
		for (uint64_t j = 0; j < nq; ++j) {
			UniBin *u = Q->UniBins + z + j; ShrBin *s = Q->ShrBins + u->six;
			memcpy(codes + off[j], u->seq, s->len);
			bud[j] = s->ed; slot[j] = (uint32_t)u->six;
		}
		bg_queries bq = {codes, off, bud, slot, (uint32_t)nq, (uint32_t)Q->numUniqQ};
		bg_hit *hits = NULL; uint64_t nh = 0;
		int rc = bg_align_batch(ctx, &bq, NULL, 0, mode, best, &hits, &nh);
		if (rc) die_gpu("bg_align_batch", rc);
		for (uint64_t h = 0; h < nh; ++h) {
			uint64_t t = hits[h].task; uint32_t clump = (uint32_t)(t / nq), j = (uint32_t)(t % nq);
			UniBin *u = Q->UniBins + z + j; ShrBin *s = Q->ShrBins + u->six;
			uint32_t refIx = clump * VECSZ + hits[h].lane;
			if (refIx >= R->totR) continue;                               /* burst.c:4444 */
			Pod p = {identity(hits[h].ed, s->len, hits[h].gap_q), refIx, hits[h].final_pos, hits[h].gap_r, hits[h].gap_q, hits[h].ed, u->rc};
			pod_push(Pods + u->six, p);
		}
		bg_free_hits(hits);
		free(off); free(bud); free(slot); free(codes);
		if (!QUIET) printf("\rSearch Progress: [%3.2f%%]", 100.0 * (double)(bound - firstQ) / (double)nqAll);
	}
	if (!QUIET) printf("\rSearch Progress: [100.00%%]\n");
	if (mode == BG_MODE_MIN) for (uint64_t i = 0; i < Q->numUniqQ; ++i) {    /* burst.c:4497-4517: drop pods above the final minimum */
		PodList *L = Pods + i; uint32_t w = 0;
		for (uint32_t k = 0; k < L->n; ++k) if (L->p[k].mismatches <= best[i]) L->p[w++] = L->p[k];
		L->n = w;
	}
	free(best);
}

/* =============================================================================================
 * Reporting (burst.c:4553-4891).  The reference's lists are push-front: iterate pods in reverse
 * discovery order.
 * ============================================================================================= */
typedef struct {
	FILE *out; Queries *Q; Refs *R; int taxasuppress;
} Rep;

static void print_row(Rep *P, uint64_t i, const Pod *rp, uint32_t rix, const char *tax) {
	Queries *Q = P->Q; Refs *R = P->R;
	uint32_t qlen = Q->ShrBins[i].len, numGap = (uint32_t)rp->numGapR + rp->numGapQ, numMis = rp->mismatches - numGap,
		alLen = qlen + numGap, mOff = R->RefStart ? R->RefStart[rix] : 0,
		stIxR = rp->finalPos - qlen + rp->numGapR + mOff, edIxR = rp->finalPos + mOff;
	if (rp->rc) { uint32_t t = stIxR; stIxR = edIxR; edIxR = t; }
	float pct = rp->score * 100;
	for (uint64_t j = Q->Offset[i]; j < Q->Offset[i + 1]; ++j) {
		if (tax) fprintf(P->out, "%s\t%s\t%f\t%u\t%u\t%u\t%u\t%u\t%d\t%u\t%u\t%" PRIu64 "\t%s\n", Q->QHead[j], R->RefHead[rix], pct,
			alLen, numMis, numGap, 1, qlen, (int)stIxR, edIxR, rp->mismatches, i, tax);
		else fprintf(P->out, "%s\t%s\t%f\t%u\t%u\t%u\t%u\t%u\t%d\t%u\t%u\t%" PRIu64 "\n", Q->QHead[j], R->RefHead[rix], pct,
			alLen, numMis, numGap, 1, qlen, (int)stIxR, edIxR, rp->mismatches, i);
	}
}

/* "first seen wins" suppression of hits on the same header whose starts are within qlen/2 (burst.c:4563-4570) */
typedef struct { uint32_t *ref, *st; uint64_t n, cap; } DupeSet;
static int dupe_hunt(DupeSet *D, Refs *R, const Pod *rp, uint32_t qlen, uint32_t rix) {
	uint32_t mOff = R->RefStart ? R->RefStart[rix] : 0, ql2 = qlen >> 1;
	uint32_t stIxR = rp->rc ? rp->finalPos + mOff : rp->finalPos - qlen + rp->numGapR + mOff;
	uint32_t mapped = R->RefMap ? R->RefMap[rix] : rix;
	for (uint64_t d = 0; d < D->n; ++d)
		if (D->ref[d] == mapped && D->st[d] + ql2 > stIxR && D->st[d] < stIxR + ql2) return 1;
	if (D->n == D->cap) { D->cap = D->cap ? D->cap * 2 : 64; D->ref = xrealloc(D->ref, D->cap * 4); D->st = xrealloc(D->st, D->cap * 4); }
	D->ref[D->n] = mapped; D->st[D->n++] = stIxR;
	return 0;
}

static const char *suppress_tax(char *buf, const char *tt, float score, uint32_t lv_limit, int use_limit) {
	/* burst.c:4874-4885 / 4820-4828: cut the taxonomy string at the level the identity supports */
	uint32_t lm, s = 0;
	if (use_limit) { for (lm = 0; lm < lv_limit && TAXLEVELS[lm] < score; ++lm); }
	else for (lm = 0; lm < 8 && TAXLEVELS[lm] < score; ++lm);
	if (!lm) return NULLTAX;
	strcpy(buf, tt);
	if (!use_limit || lm < lv_limit) for (int x = 0; buf[x]; ++x) if (buf[x] == ';' && ++s == lm) { buf[x] = 0; break; }
	return buf;
}

/* The rows of a query depend on nothing but its own pods (and, for CAPITALIST, on reference counts tallied beforehand), so the team
 * formats blocks of queries into memory streams (printf's %f is most of the reporting time: 0.6 s per million rows on one thread) and
 * the blocks are written out in query order: the same bytes as the sequential loop. */
typedef void (*RangeFn)(Rep *P, PodList *Pods, uint64_t i0, uint64_t i1, void *shared);
static void report_in_blocks(Rep *P, PodList *Pods, RangeFn fn, void *shared) {
	Queries *Q = P->Q;
	const uint64_t CH = getenv("BURST_B200_REPORT_BLOCK") && atoi(getenv("BURST_B200_REPORT_BLOCK")) > 0 ? (uint64_t)atoi(getenv("BURST_B200_REPORT_BLOCK")) : 8192;   /* queries per block (the variable is for the tests) */
	const uint64_t nch = (Q->numUniqQ + CH - 1) / CH;
	int nthr = THREADS < 1 ? 1 : THREADS;
	if (nthr == 1 || nch < 2) { fn(P, Pods, 0, Q->numUniqQ, shared); return; }
	const uint64_t WAVE = (uint64_t)nthr * 8;                               /* blocks in memory at a time */
	char **blk = xcalloc(WAVE, sizeof(*blk)); size_t *len = xcalloc(WAVE, sizeof(*len));
	for (uint64_t c0 = 0; c0 < nch; c0 += WAVE) {
		const uint64_t c1 = MIN(nch, c0 + WAVE);
		int failed = 0;
		#pragma omp parallel for schedule(dynamic, 1) num_threads(nthr)
		for (uint64_t c = c0; c < c1; ++c) {
			Rep P2 = *P; blk[c - c0] = NULL; len[c - c0] = 0;
			P2.out = open_memstream(&blk[c - c0], &len[c - c0]);
			if (!P2.out) { failed = 1; continue; }
			fn(&P2, Pods, c * CH, MIN(Q->numUniqQ, (c + 1) * CH), shared);
			fclose(P2.out);
		}
		if (failed) { fputs("OOM: report buffers\n", stderr); exit(3); }
		for (uint64_t c = c0; c < c1; ++c) { if (len[c - c0]) fwrite(blk[c - c0], 1, len[c - c0], P->out); free(blk[c - c0]); }
	}
	free(blk); free(len);
}

/* rows of queries [i0, i1) in order (burst.c:4847-4891) */
static void report_best_range(Rep *P, PodList *Pods, uint64_t i0, uint64_t i1, void *shared) {
	Refs *R = P->R; (void)shared;
	char *buf = xmalloc(1 << 20);
	for (uint64_t i = i0; i < i1; ++i) {
		PodList *L = Pods + i; if (!L->n) continue;
		const Pod *best = &L->p[L->n - 1];
		for (int64_t k = (int64_t)L->n - 2; k >= 0; --k) {
			const Pod *rp = &L->p[k];
			if (rp->mismatches < best->mismatches || (rp->mismatches == best->mismatches && rp->score > best->score) ||
			    (rp->mismatches == best->mismatches && rp->score == best->score && R->RefIxSrt[rp->refIx] < R->RefIxSrt[best->refIx])) best = rp;
		}
		uint32_t rix = R->RefIxSrt[best->refIx];
		const char *tax = NULL;
		if (taxa_parsed) { tax = taxa_lookup(R->RefHead[rix]); if (P->taxasuppress) tax = suppress_tax(buf, tax, best->score, 0, 0); }
		print_row(P, i, best, rix, tax);
	}
	free(buf);
}
static void report_best(Rep *P, PodList *Pods) { report_in_blocks(P, Pods, report_best_range, NULL); }

typedef struct { const Pod *rp; uint32_t rix; } RowRef;
/* expand a pod over the de-duplicated originals it stands for (burst.c:4601-4616) */
#define FOR_EACH_RIX(R, rp, rixvar, ...) \
	if ((R)->RefDedupIx) { for (uint32_t k_ = (R)->RefDedupIx[(rp)->refIx]; k_ < (R)->RefDedupIx[(rp)->refIx + 1]; ++k_) { uint32_t rixvar = (R)->TmpRIX[k_]; __VA_ARGS__ } } \
	else { uint32_t rixvar = (R)->RefIxSrt[(rp)->refIx]; __VA_ARGS__ }

static void report_allpaths_range(Rep *P, PodList *Pods, uint64_t i0, uint64_t i1, void *shared) {   /* burst.c:4582-4692 */
	Queries *Q = P->Q; Refs *R = P->R; const int forage = *(const int *)shared;
	DupeSet D = {0}; RowRef *rows = NULL; uint64_t rcap = 0;
	for (uint64_t i = i0; i < i1; ++i) {
		PodList *L = Pods + i; if (!L->n) continue;
		uint32_t qlen = Q->ShrBins[i].len, bm = 255; uint64_t nrows = 0; D.n = 0;
		const Pod *best = &L->p[L->n - 1];
		for (int64_t k = (int64_t)L->n - 2; k >= 0; --k) if (L->p[k].mismatches < best->mismatches) best = &L->p[k];
		bm = best->mismatches;
		if (!forage && !(best->score != 0)) continue;                       /* burst.c:4598 */
		for (int64_t k = (int64_t)L->n - 1; k >= 0; --k) {
			const Pod *rp = &L->p[k];
			if (!forage && rp->mismatches != bm) continue;
			FOR_EACH_RIX(R, rp, rix, {
				if (!dupe_hunt(&D, R, rp, qlen, rix)) {
					if (nrows == rcap) { rcap = rcap ? rcap * 2 : 64; rows = xrealloc(rows, rcap * sizeof(*rows)); }
					rows[nrows].rp = rp; rows[nrows++].rix = rix;
				}
			})
		}
		for (uint64_t j = Q->Offset[i]; j < Q->Offset[i + 1]; ++j) for (uint64_t z = 0; z < nrows; ++z) {
			/* one row per (duplicate, hit): print_row prints all duplicates, so emit per duplicate here */
			const Pod *rp = rows[z].rp; uint32_t rix = rows[z].rix;
			uint32_t numGap = (uint32_t)rp->numGapR + rp->numGapQ, numMis = rp->mismatches - numGap, alLen = qlen + numGap,
				mOff = R->RefStart ? R->RefStart[rix] : 0, a = rp->finalPos - qlen + rp->numGapR + mOff, b = rp->finalPos + mOff,
				stIxR = rp->rc ? b : a, edIxR = rp->rc ? a : b;
			float pct = rp->score * 100;
			if (taxa_parsed) fprintf(P->out, "%s\t%s\t%f\t%u\t%u\t%u\t%u\t%u\t%d\t%u\t%u\t%" PRIu64 "\t%s\n", Q->QHead[j], R->RefHead[rix], pct,
				alLen, numMis, numGap, 1, qlen, (int)stIxR, edIxR, rp->mismatches, i, taxa_lookup(R->RefHead[rix]));
			else fprintf(P->out, "%s\t%s\t%f\t%u\t%u\t%u\t%u\t%u\t%d\t%u\t%u\t%" PRIu64 "\n", Q->QHead[j], R->RefHead[rix], pct,
				alLen, numMis, numGap, 1, qlen, (int)stIxR, edIxR, rp->mismatches, i);
		}
	}
	free(rows); free(D.ref); free(D.st);
}
static void report_allpaths_or_forage(Rep *P, PodList *Pods, int forage) { report_in_blocks(P, Pods, report_allpaths_range, &forage); }

static int cmp_str(const void *a, const void *b) { return strcmp(*(char *const *)a, *(char *const *)b); }

typedef struct { const size_t *RefCounts; const uint32_t *start; } CapShared;
static void report_capitalist_range(Rep *P, PodList *Pods, uint64_t i0, uint64_t i1, void *shared);
static void report_capitalist(Rep *P, PodList *Pods) {                  /* burst.c:4694-4846 */
	Queries *Q = P->Q; Refs *R = P->R;
	uint32_t maxIX = 0;
	for (uint32_t i = 0; i < R->totR; ++i) if (R->RefIxSrt[i] > maxIX) maxIX = R->RefIxSrt[i];
	uint64_t numBins = (uint64_t)maxIX + 1;
	if (R->RefMap) for (uint32_t i = 0; i < R->origTotR; ++i) if ((uint64_t)R->RefMap[i] + 1 > numBins) numBins = (uint64_t)R->RefMap[i] + 1;
	size_t *RefCounts = xcalloc(numBins, sizeof(*RefCounts)), tot = 0;
	DupeSet D = {0};
	uint32_t *start = xcalloc(Q->numUniqQ, sizeof(*start));      /* index (in list order) of the first minimum pod */
	/* pass 1+2 (4700-4727): find the first minimum in list order, tally references over the pods from it on */
	for (uint64_t i = 0; i < Q->numUniqQ; ++i) {
		PodList *L = Pods + i; if (!L->n) continue;
		int64_t bk = (int64_t)L->n - 1;
		for (int64_t k = (int64_t)L->n - 2; k >= 0; --k) if (L->p[k].mismatches < L->p[bk].mismatches) bk = k;
		start[i] = (uint32_t)bk; D.n = 0;
		uint32_t qlen = Q->ShrBins[i].len, bm = L->p[bk].mismatches;
		for (int64_t k = bk; k >= 0; --k) {
			const Pod *rp = &L->p[k];
			if (rp->mismatches != bm) continue;
			FOR_EACH_RIX(R, rp, rix, {
				if (!dupe_hunt(&D, R, rp, qlen, rix)) { ++RefCounts[R->RefMap ? R->RefMap[rix] : rix]; ++tot; }
			})
		}
	}
	printf("CAPITALIST: Processed %zu investments\n", tot);
	free(D.ref); free(D.st);
	CapShared CS = {RefCounts, start};
	report_in_blocks(P, Pods, report_capitalist_range, &CS);             /* pass 3: a query's row needs its own pods and the finished counts only */
	free(RefCounts); free(start);
}
static void report_capitalist_range(Rep *P, PodList *Pods, uint64_t i0, uint64_t i1, void *shared) {
	Queries *Q = P->Q; Refs *R = P->R;
	const size_t *RefCounts = ((CapShared *)shared)->RefCounts; const uint32_t *start = ((CapShared *)shared)->start;
	DupeSet D = {0};
	char **Taxa = NULL; uint32_t *Div = NULL; char *Taxon = NULL; uint64_t tcap = 0;
	if (taxa_parsed) Taxon = xmalloc(1000000);
	for (uint64_t i = i0; i < i1; ++i) {
		PodList *L = Pods + i; if (!L->n) continue;
		const Pod *first = &L->p[start[i]], *best = first;
		uint32_t tix = 0, bestmap = 0, bestrix = R->RefIxSrt[first->refIx], qlen = Q->ShrBins[i].len; float best_score = -1.f;
		D.n = 0;
		for (int64_t k = start[i]; k >= 0; --k) {
			const Pod *rp = &L->p[k];
			if (rp->mismatches > first->mismatches) continue;
			FOR_EACH_RIX(R, rp, rix, {
				if (!dupe_hunt(&D, R, rp, qlen, rix)) {
					uint32_t mapped = R->RefMap ? R->RefMap[rix] : rix;
					if (taxa_parsed) {
						if (tix == tcap) { tcap = tcap ? tcap * 2 : 64; Taxa = xrealloc(Taxa, tcap * sizeof(*Taxa)); Div = xrealloc(Div, tcap * sizeof(*Div)); }
						Taxa[tix++] = taxa_lookup(R->RefHead[rix]);
						if (rp->score > best_score) best_score = rp->score;
					}
					/* burst.c:4763-4765: the pod currently held is overridden by its own later originals;
					 * another pod takes over on a larger global count, ties to the smaller header index */
					if (best == rp || RefCounts[mapped] > RefCounts[bestmap] || (RefCounts[mapped] == RefCounts[bestmap] && mapped < bestmap))
						best = rp, bestmap = mapped, bestrix = rix;
				}
			})
		}
		const char *FinalTaxon = NULL;
		if (taxa_parsed) {                                              /* LCA interpolation, burst.c:4781-4829 */
			uint32_t lv = (uint32_t)-1;
			if (tix == 1) { strcpy(Taxon, Taxa[0]); FinalTaxon = Taxon; }
			else {
				qsort(Taxa, tix, sizeof(*Taxa), cmp_str);
				uint32_t maxDiv = 0; Div[0] = 0;
				for (uint32_t z = 1; z < tix; ++z) {
					uint32_t x; Div[z] = 0;
					for (x = 0; Taxa[z - 1][x] && Taxa[z - 1][x] == Taxa[z][x]; ++x) Div[z] += Taxa[z][x] == ';';
					Div[z] += !Taxa[z - 1][x];
					if (Div[z] > maxDiv) maxDiv = Div[z];
				}
				if (!maxDiv) { Taxon[0] = 0; FinalTaxon = Taxon; }
				else {
					uint32_t cutoff = tix - tix / TAXACUT, st = 0, ed = tix;
					for (lv = 1; lv <= maxDiv; ++lv) {
						uint32_t accum = 1;
						for (uint32_t z = st + 1; z < ed; ++z) {
							if (Div[z] >= lv) ++accum;
							else if (accum >= cutoff) { ed = z; break; }
							else accum = 1, st = z;
						}
						if (accum < cutoff) break;
						cutoff = accum - accum / TAXACUT;
					}
					uint32_t s = 0;
					if (ed) --ed;
					--lv;
					for (st = 0; Taxa[ed][st] && (s += Taxa[ed][st] == ';') < lv; ++st) Taxon[st] = Taxa[ed][st];
					Taxon[st] = 0; FinalTaxon = Taxon;
				}
			}
			if (P->taxasuppress) {
				uint32_t lm, s = 0;
				for (lm = 0; lm < lv && lm < 8 && TAXLEVELS[lm] < best_score; ++lm);
				if (!lm) FinalTaxon = NULLTAX;
				else if (lm < lv) for (int x = 0; Taxon[x]; ++x) if (Taxon[x] == ';' && ++s == lm) { Taxon[x] = 0; break; }
			}
		}
		print_row(P, i, best, bestrix, FinalTaxon);
	}
	free(D.ref); free(D.st); free(Taxa); free(Div); free(Taxon);
}

/* =============================================================================================
 * main: the reference's command line (burst.c:4902-5103)
 * ============================================================================================= */
static void usage(void) {
	printf("\nBURST aligner, B200 build (" VER ")\n");
	puts("Same command line as BURST (burst.c:102-150) for the alignment path:");
	puts("--references (-r) <name>: FASTA/edx DB of reference sequences [required]");
	puts("--accelerator (-a) <name>: uses a helper DB (acx) [optional]");
	puts("--queries (-q) <name>: FASTA file of queries to search [required]");
	puts("--output (-o) <name>: Blast6 file for output alignments [required]");
	puts("--forwardreverse (-fr), --whitespace (-w), --nwildcard (-y), --npenalize (-n)");
	puts("--taxonomy (-b) <name>, --taxacut (-bc) <num>, --taxa_ncbi (-bn), --taxasuppress (-bs) [STRICT]");
	puts("--mode (-m) BEST | ALLPATHS | CAPITALIST [default] | FORAGE");
	puts("--id (-i) <decimal> [0.97], --threads (-t) <int> (host threads for candidate generation), --skipambig (-sa), --heuristic (-hr)");
	puts("--gpu <int>: first CUDA device to use [0];  --gpus <int>: number of devices (queries are sharded, DB replicated) [1];  --noprogress");
	puts("--makedb (-d) [DNA|RNA|QUICK] [qLen]: write -o <edx> (and -a <acx>, word length --acx-n 12|15 [12]) from -r <fasta>; -s [len] shears");
	exit(1);
}
static int is_edx(const char *fn) {                                  /* burst.c:4894-4901 */
	FILE *f = fopen(fn, "rb");
	if (!f) { fputs("ERROR: invalid input file.\n", stderr); exit(1); }
	int c = fgetc(f); fclose(f);
	if (c == EOF) { fputs("ERROR: invalid input file.\n", stderr); exit(1); }
	return (uint8_t)c >> 7;
}

typedef struct { bg_ctx **ctxs; int n, dev0; const uint8_t *S; int rc; char msg[256]; } EngineStart;
static void *engine_start(void *p) {                                    /* --gpus N: devices dev0 .. dev0+N-1, one context each */
	EngineStart *E = p;
	for (int g = 0; g < E->n; ++g) {
		E->ctxs[g] = NULL;
		int rc = bg_init(E->dev0 + g, &E->ctxs[g]);
		if (!rc) rc = bg_set_scoring(E->ctxs[g], E->S);
		if (rc) { E->rc = rc; snprintf(E->msg, sizeof(E->msg), "%s", bg_last_error()); return NULL; }
	}
	return NULL;
}
int main(int argc, char *argv[]) {
	Queries Q; memset(&Q, 0, sizeof(Q));
	Refs R; char *ref_FN = 0, *query_FN = 0, *output_FN = 0, *xcel_FN = 0, *tax_FN = 0; int taxasuppress = 0, makedb = 0;
	printf("This is BURST [" VER "]\n");
	if (argc < 2) usage();
#define NEEDARG(msg) if (++i == argc || argv[i][0] == '-') { puts("ERROR: " msg); exit(1); }
#define OPT(l, s) (!strcmp(argv[i], l) || !strcmp(argv[i], s))
	for (int i = 1; i < argc; ++i) {
		if (OPT("--references", "-r")) { NEEDARG("--references requires filename argument") ref_FN = argv[i]; }
		else if (OPT("--queries", "-q")) { NEEDARG("--queries requires filename argument") query_FN = argv[i]; }
		else if (OPT("--output", "-o")) { NEEDARG("--output requires filename argument") output_FN = argv[i]; }
		else if (OPT("--forwardreverse", "-fr")) { Q.rc = 1; printf(" --> Also considering the reverse complement of reads\n"); }
		else if (OPT("--whitespace", "-w")) { Q.incl_whitespace = 1; printf(" --> Allowing whitespace in query name output\n"); }
		else if (OPT("--npenalize", "-n")) { Z = 1; printf(" --> Setting N penalty (ref N vs query A/C/G/T)\n"); }
		else if (OPT("--nwildcard", "-y")) { Z = 0; printf(" --> Setting N's and X's to wildcards (match anything)\n"); }
		else if (OPT("--xalphabet", "-x")) { fputs("ERROR: --xalphabet is not supported by this build (the reference itself aborts on it, see DESIGN.md)\n", stderr); exit(1); }
		else if (OPT("--taxonomy", "-b")) { NEEDARG("--taxonomy requires filename argument") tax_FN = argv[i]; printf(" --> Assigning taxonomy based on mapping file: %s\n", tax_FN); }
		else if (OPT("--mode", "-m")) {
			NEEDARG("--mode requires an argument (see -h)")
			if (!strcmp(argv[i], "BEST")) RUNMODE = BEST; else if (!strcmp(argv[i], "ALLPATHS")) RUNMODE = ALLPATHS;
			else if (!strcmp(argv[i], "CAPITALIST")) RUNMODE = CAPITALIST; else if (!strcmp(argv[i], "FORAGE")) RUNMODE = FORAGE;
			else if (!strcmp(argv[i], "ANY")) { fputs("ERROR: -m ANY (first hit in thread-arrival order) is not supported by this build\n", stderr); exit(1); }
			else if (!strcmp(argv[i], "MATRIX")) { fputs("ERROR: Matrix mode is no longer supported\n", stderr); exit(1); }
			else { printf("Unsupported run mode '%s'\n", argv[i]); exit(1); }
			printf(" --> Setting run mode to %s\n", argv[i]);
		}
		else if (OPT("--makedb", "-d")) {                                  /* burst.c:4969-4985 */
			makedb = 1; const char *dbsel = "QUICK";
			if (i + 1 != argc && argv[i + 1][0] != '-' && !atol(argv[i + 1])) {
				++i;
				if (strcmp(argv[i], "DNA") && strcmp(argv[i], "RNA") && strcmp(argv[i], "QUICK")) { printf("Unsupported makedb mode '%s'\n", argv[i]); exit(1); }
				dbsel = argv[i];
			}
			if (i + 1 != argc && argv[i + 1][0] != '-') {
				DB_QLEN = atol(argv[++i]);
				if (DB_QLEN <= 0) { fprintf(stderr, "ERROR: bad max query length '%s'\n", argv[i]); exit(1); }
			}
			printf(" --> Creating %s database (assuming max query length %ld)\n", dbsel, DB_QLEN);
		}
		else if (!strcmp(argv[i], "--acx-n")) { NEEDARG("--acx-n requires 12 or 15") ACX_N = atoi(argv[i]); if (ACX_N != 12 && ACX_N != 15) { fputs("ERROR: --acx-n must be 12 or 15\n", stderr); exit(1); } }
		else if (OPT("--accelerator", "-a")) { NEEDARG("--accelerator requires filename argument") xcel_FN = argv[i]; DO_ACCEL = 1; printf(" --> Using accelerator file %s\n", xcel_FN); }
		else if (OPT("--taxacut", "-bc")) {
			NEEDARG("--taxacut requires numeric argument")
			int t = atoi(argv[i]);
			if (t < 2) { double fl = 1.0 / (1.0 - atof(argv[i])); t = (int)(fl + 0.5); printf(" --> Taxacut: converting %s to %d...\n", argv[i], t); }
			if (t < 2) { fputs("ERROR: taxacut must be >= 2\n", stderr); exit(1); }
			TAXACUT = (uint32_t)t; printf(" --> Ignoring 1/%u disagreeing taxonomy calls\n", TAXACUT);
		}
		else if (OPT("--taxa_ncbi", "-bn")) { TAXA_NCBI = 1; printf(" --> Using NCBI header formatting for taxonomy lookups\n"); }
		else if (OPT("--skipambig", "-sa")) { Q.skipAmbig = 1; printf(" --> Skipping highly ambiguous sequences\n"); }
		else if (OPT("--taxasuppress", "-bs")) {
			taxasuppress = 1;
			if (i + 1 != argc && argv[i + 1][0] != '-') { if (!strcmp(argv[++i], "STRICT")) TAXLEVELS = TAXLEVELS_STRICT; else { fprintf(stderr, "ERROR: Unrecognized taxasuppress '%s'\n", argv[i]); exit(1); } }
			printf(" --> Surpressing taxonomic specificity by alignment identity%s\n", TAXLEVELS == TAXLEVELS_STRICT ? " [STRICT]" : "");
		}
		else if (OPT("--id", "-i")) {
			NEEDARG("--id requires decimal argument")
			THRES = (float)atof(argv[i]);
			if (THRES > 1.f || THRES < 0.f) { puts("Invalid id range [0-1]"); exit(1); }
			if (THRES < 0.01f) THRES = 0.01f;
			printf(" --> Setting identity threshold to %f\n", THRES);
		}
		else if (OPT("--threads", "-t")) { NEEDARG("--threads requires integer argument") THREADS = atoi(argv[i]) > 0 ? atoi(argv[i]) : 1; printf(" --> Setting threads to %d (bunch size and host-side candidate generation; the DP runs on the GPU)\n", THREADS); }
		else if (OPT("--shear", "-s")) { REBASE = 1; if (i + 1 != argc && argv[i + 1][0] != '-') REBASE_AMT = atol(argv[++i]); if (!REBASE_AMT) REBASE = 0; }
		else if (OPT("--heuristic", "-hr")) { DO_HEUR = 1; printf(" --> WARNING: Heuristic mode set; optimality not guaranteed at low ids\n"); }
		else if (!strcmp(argv[i], "--noprogress")) { QUIET = 1; printf(" --> Surpressing progress indicator\n"); }
		else if (!strcmp(argv[i], "--gpu")) { NEEDARG("--gpu requires integer argument") GPU_DEVICE = atoi(argv[i]); }
		else if (!strcmp(argv[i], "--shard-refs")) { SHARD_REFS = 1; printf(" --> Sharding the reference database over the GPUs (all-reduce MIN on the per-read minima)\n"); }
		else if (!strcmp(argv[i], "--device-candidates")) { DEVICE_CAND = 1; printf(" --> Candidate generation on the GPU (accelerator resident in device memory)\n"); }
		else if (!strcmp(argv[i], "--gpus")) { NEEDARG("--gpus requires integer argument") NGPU = atoi(argv[i]); if (NGPU < 1 || NGPU > 64) { fputs("ERROR: --gpus must be 1..64\n", stderr); exit(1); } }
		else if (OPT("--fingerprint", "-f") || OPT("--prepass", "-p") || OPT("--unique", "-u")) { fprintf(stderr, "ERROR: %s selects a heuristic/legacy path that this build does not provide (see DESIGN.md, out of scope)\n", argv[i]); exit(1); }
		else if (OPT("--cache", "-c") || OPT("--latency", "-l") || OPT("--clustradius", "-cr") || OPT("--dbpartition", "-dp")) { NEEDARG("option requires integer argument") }
		else if (OPT("--help", "-h")) usage();
		else { printf("ERROR: Unrecognized command-line option: %s\n", argv[i]); puts("See help by running with just '-h'"); exit(1); }
	}
	if (makedb) {                                                       /* burst.c:5118-5134: no GPU involved */
		if (!ref_FN || !output_FN) { puts("ERROR: -r and -o are required"); exit(1); }
		if (is_edx(ref_FN)) { fputs("ERROR: DBs can't make DBs.\n", stderr); exit(1); }
		init_char2num();
		make_db(ref_FN, output_FN, xcel_FN, DB_QLEN, ACX_N, Q.skipAmbig);
		return 0;
	}
	if (!ref_FN || !query_FN || !output_FN) { puts("ERROR: -r, -q and -o are required"); exit(1); }
	FILE *output = fopen(output_FN, "wb");
	if (!output) { fprintf(stderr, "ERROR: Cannot open output: %s\n", output_FN); exit(2); }
	setvbuf(output, 0, _IOFBF, 1 << 22);
	double start = now(), tph = start;
#define PHASE(name) do { double t_ = now(); printf(" --> [time] %-28s %8.3f s\n", name, t_ - tph); tph = t_; } while (0)
	init_char2num();

	/* the engine starts (CUDA context per device: 0.5 - 3 s on a cold box) while the files are read; without a device nothing below can run */
	bg_ctx *ctxs[64]; int rc;
	uint8_t S[256]; bg_default_scoring(Z, S);
	EngineStart ES; memset(&ES, 0, sizeof(ES)); ES.ctxs = ctxs; ES.n = NGPU; ES.dev0 = GPU_DEVICE; ES.S = S;
	pthread_t es_thread;
	if (pthread_create(&es_thread, NULL, engine_start, &ES)) { fputs("ERROR: cannot start a thread\n", stderr); exit(4); }
	Acx A; memset(&A, 0, sizeof(A));
	if (DO_ACCEL) { load_acx(xcel_FN, &A); PHASE("accelerator load"); }
	int usedb = is_edx(ref_FN);
	if (usedb) { puts("\nEDB database provided. Parsing..."); load_edx(ref_FN, &R); }
	PHASE("database load");
	if (tax_FN) load_taxonomy(tax_FN);
	load_queries(query_FN, &Q);
	PHASE("query parse/sort");
	pthread_join(es_thread, NULL);
	if (ES.rc) { fprintf(stderr, "ERROR: cannot start the GPU engine: %s\n", ES.msg); exit(3); }
	bg_ctx *ctx = ctxs[0];
	PHASE("GPU engine start (rest)");
	if (!usedb) load_fasta_refs(ref_FN, &R);
	else if (R.shear && (uint32_t)(Q.maxLenQ / THRES) > R.shear) {
		fputs("ERROR: DB incompatible with selected queries/identity.\n", stderr);
		if (!DO_HEUR) exit(1);
		fputs("!!! WARNING: Error overridden by use of heuristic mode!\n", stderr);
	}
	if (SHARD_REFS && NGPU > 1) {                                        /* contiguous clump ranges of about equal bytes, one per GPU */
#ifndef BURST_NCCL
		fputs("ERROR: this build has no NCCL: --shard-refs is unavailable\n", stderr); exit(1);
#endif
		if (!DO_ACCEL) { fputs("ERROR: --shard-refs needs an accelerator (-a)\n", stderr); exit(1); }
		uint64_t *boff = xmalloc(((size_t)R.numRclumps + 1) * 8); boff[0] = 0;
		for (uint32_t i = 0; i < R.numRclumps; ++i) boff[i + 1] = boff[i] + (uint64_t)((R.ClumpLen[i] + 1) / 2) * 16;
		uint32_t lo = 0;
		for (int g = 0; g < NGPU; ++g) {
			uint32_t hi = lo;
			uint64_t want = boff[R.numRclumps] / (uint64_t)NGPU * (uint64_t)(g + 1);
			while (hi < R.numRclumps && (g == NGPU - 1 || boff[hi] < want)) ++hi;
			if (hi == lo) { fputs("ERROR: more GPUs than clumps\n", stderr); exit(1); }
			printf(" --> shard %d: clumps [%u, %u), %.1f MB\n", g, lo, hi, (double)(boff[hi] - boff[lo]) / 1e6);
			if ((rc = bg_load_db(ctxs[g], R.packed + boff[lo], R.ClumpLen + lo, hi - lo, lo))) die_gpu("bg_load_db", rc);
			lo = hi;
		}
		free(boff);
	} else
	for (int g = 0; g < NGPU; ++g) if ((rc = bg_load_db(ctxs[g], R.packed, R.ClumpLen, R.numRclumps, 0))) die_gpu("bg_load_db", rc);   /* queries are sharded, the database is replicated */

	PHASE("database to GPU");
	PodList *Pods = xcalloc(Q.numUniqQ, sizeof(*Pods));
	int mode = RUNMODE == FORAGE ? BG_MODE_ALL : BG_MODE_MIN;
#ifdef BURST_NCCL
	if (DO_ACCEL && SHARD_REFS && NGPU > 1) accel_search_sharded(ctxs, NGPU, &Q, &R, &A, Pods, mode, THREADS); else
#endif
	if (DO_ACCEL && DEVICE_CAND && Q.QBins[0] == 0 && Q.maxLenQ <= 512 + (uint32_t)SCOUR_N - 1) accel_search_device(ctxs, NGPU, &Q, &R, &A, Pods, mode, THREADS);
	else if (DO_ACCEL) {
		if (DEVICE_CAND) printf(" --> [Accel] --device-candidates not used: %s\n", Q.QBins[0] ? "queries with ambiguous bases need the host's variant expansion" : "queries longer than the device word tables hold");
		accel_search(ctxs, NGPU, &Q, &R, &A, Pods, mode, THREADS);
	}
	/* queries the accelerator cannot vouch for (or all of them without -a) go all-vs-all, burst.c:4320-4323 */
	uint64_t firstQ = DO_ACCEL ? Q.QBins[1] : 0;
	if (firstQ != Q.newUniqQ && !(DO_ACCEL && Q.skipAmbig)) {
		if (SHARD_REFS && NGPU > 1) { fputs("ERROR: some queries need the all-vs-all search, which --shard-refs does not cover; use -sa to skip them\n", stderr); exit(1); }
		search_all_vs_all(ctx, &Q, &R, firstQ, Pods, mode);
	}
	PHASE("search");
	printf("Search complete. Consolidating results...\n");
	Rep P = {output, &Q, &R, taxasuppress};
	if (RUNMODE == BEST) report_best(&P, Pods);
	else if (RUNMODE == ALLPATHS) report_allpaths_or_forage(&P, Pods, 0);
	else if (RUNMODE == FORAGE) report_allpaths_or_forage(&P, Pods, 1);
	else report_capitalist(&P, Pods);
	fclose(output);
	PHASE("report");
	for (int g = 0; g < NGPU; ++g) bg_free(ctxs[g]);
	printf("\nAlignment time: %f seconds\n", now() - start);
	return 0;
}
