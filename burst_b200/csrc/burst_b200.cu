// burst_b200.cu -- sm_100a kernels + C ABI (include/burst_b200.h) of the B200 alignment engine.
//
// What the reference does per (query, clump) pair:
//   pass 1  aded_mat16L / aded_mat16   burst.c:1003-1204   16-lane banded edit distance
//   pass 2  reScoreM_mat16             burst.c:713-886     (score, shift, shiftR) + end column
// and how this file restructures it for the GPU (DESIGN.md section 2 has the full argument):
//   k_qinfo/k_qprep            per query: budget/length record, nibble-packed bases (k_filter builds its Myers match vectors itself)
//   k_seed         a block walks its RUNS (run = one clump visit by <= 16 consecutive queries, the reference's
//                  "unpack the clump once, then loop over the bunch", burst.c:4141-4157) in rounds, one run per group
//                  of 16 threads, one thread per reference lane.  Pigeonhole: an alignment with <= k errors leaves
//                  one of k+1 disjoint query stretches unchanged, so some word-aligned (or half-word-aligned) window
//                  of w reference bases equals one of `stride` windows at the end of that stretch.  The bunch's
//                  query windows go into a blocked Bloom filter + hash table in shared memory (built once per bunch,
//                  shared by its runs); each lane is streamed once, one hash probe per 8 (or 4) columns; a thread
//                  verifies its own Bloom hits exactly through the table -> seeds -> disjoint diagonal clusters
//                  [d-k, d+k].  Clumps arrive by 128-bit loads or, optionally, 1-D TMA bulk copies + mbarriers.
//   k_filter       queries k_seed cannot take (many errors / short stretches / IUPAC bases): Myers/Hyyro bit-vector
//                  semi-global DP of the query's first P <= 32 rows; columns with row-P value <= k
//                  are seeds (every <= k alignment has a <= k prefix); hull of the seed diagonals.
//   k_extend       one thread per surviving (task, lane, cluster): exact banded DP, carrying the
//                  reference's pass-2 triple packed in one 32-bit key so that diag/up/left selection
//                  with its tie-break order is a 3-operand add-min (DPX VIADDMNMX).  Yields pass 1's
//                  distance and pass 2's (numGapQ, numGapR, finalPos) in one sweep; atomicMin keeps
//                  the per-slot best (ShrBins[].ed, burst.c:4220).
//   k_select       merges the clusters of a lane and keeps the lanes the reference keeps
//                  (burst.c:4219-4229, 4497).
// Values <= budget are exact and identical to the reference's saturating u8 arithmetic because
// every cell > maxED is treated as absent there too (burst.c:1053-1054, 802-803).
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <algorithm>
#include <vector>
#include <mutex>
#include <memory>
#include <chrono>
#include <type_traits>
#include "burst_b200.h"

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...) {
	va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
	return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return fail(BG_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)

extern "C" const char *bg_last_error(void) { return g_err; }

// ---------------------------------------------------------------------------------------------
// device-side data
// ---------------------------------------------------------------------------------------------
// Prefix rows of the Myers filter, coded in one byte: up to 32 rows = one word holding exactly that many; else 32 + NW for NW in
// {2, 4, 8, 16, 32} words of 32 rows.  NW grows with the budget (about 4 k rows) as long as the query is that long.
__host__ __device__ __forceinline__ uint32_t filt_words(uint32_t P) { return P <= 32 ? 1u : P - 32u; }
__host__ __device__ __forceinline__ uint32_t filt_rows(uint32_t P) { return P <= 32 ? P : (P - 32u) * 32u; }
__host__ __device__ __forceinline__ uint8_t filt_code(uint64_t len, uint32_t k) {
	uint32_t nw = 1; const uint32_t want = (k + 7) / 8;
	while (nw < 32 && nw < want && (uint64_t)nw * 64 <= len) nw <<= 1;
	return (uint8_t)(nw == 1 ? (len < 32 ? len : 32) : 32 + nw);
}
struct QInfo {            // one query of the batch
	uint64_t off;         // into codes
	uint32_t len;
	uint32_t slot;
	uint16_t k;           // budget (Emac)
	uint8_t  P;           // rows covered by the Myers prefix filter, coded (filt_code)
	uint8_t  cls;         // bit 0: handled by k_seed (else by k_filter, Myers); bit 1: every base is a plain A/C/G/T
};
struct Surv {             // one diagonal cluster of a (task, lane) that survived the filter
	uint32_t task;        // run * 16 + query-in-run (batch-wide)
	int32_t  lo;          // lowest diagonal (x - y) of the band
	uint32_t w_lane;      // band width << 8 | clusters-in-group << 4 (first of a group, else 0) | lane
	uint32_t scratch;     // offset into the global band scratch (generic kernel only)
};
struct Res { uint32_t a, b, slot; };   // a = ed | gap_q << 8 | gap_r << 16 | valid << 31 ; b = final_pos ; slot of the query

// Where the work list comes from: explicit runs, or all-vs-all tiles (run r = clump r / ntiles,
// queries 16 * (r % ntiles) ..).
// q_base / run_base: a slice of a larger batch (pipelined upload) holds queries q_base.. and runs run_base..;
// run records keep their batch-wide query numbers, task ids carried by survivors and hits are batch-wide.
struct Work {
	const bg_run *runs; uint64_t nruns; uint32_t nq, ntiles, first_clump, num_clumps, q_base, run_base;
};
__device__ __forceinline__ bool get_run(const Work &W, uint64_t r, uint32_t &c, uint32_t &q0, uint32_t &n) {
	if (W.runs) { const bg_run R = W.runs[r]; c = R.clump; q0 = R.query0 - W.q_base; n = R.nq; }
	else { c = (uint32_t)(r / W.ntiles) + W.first_clump; q0 = (uint32_t)(r % W.ntiles) * BG_RUN_MAX; n = min((uint32_t)BG_RUN_MAX, W.nq - q0); }
	c -= W.first_clump;
	return c < W.num_clumps;                      // other shards' clumps are skipped
}

// counters: [0] survivors, [1] scratch words, [2] hits, [3] error flag
enum { C_SURV = 0, C_SCRATCH = 1, C_HITS = 2, C_ERR = 3 };

// The pass-2 cell (score, shift, shiftR) of burst.c:763-799 as ONE ordered key:
//   bits 31..22 score   (min wins)
//   bits 21..11 2047 - shift   (on equal score the larger shift wins, burst.c:776-777/794-795)
//   bits 10..9  which predecessor: 0 diag, 1 up, 2 left (on a full tie the earlier one in the
//               reference's fixed order diag -> up -> left keeps the cell)
//   bits  8..0  shiftR  (carried along with the winner, never compared)
// Stored cells have the predecessor bits cleared.
#define KEY_ZERO   (2047u << 11)
#define KEY_UP     ((1u << 22) + (1u << 9) + 1u)             // score+1, via up, shiftR+1
#define KEY_LEFT   ((1u << 22) - (1u << 11) + (2u << 9))     // score+1, shift+1, via left
#define KEY_CLEAR  (~(3u << 9))
#define KEY_NONE   0xFFFFFFFFu

__device__ __forceinline__ uint32_t key_col0(uint32_t y) { return (y << 22) | KEY_ZERO | y; }  // burst.c:747-750

__device__ __forceinline__ uint32_t viaddmin(uint32_t a, uint32_t b, uint32_t c) {   // min(a + b, c)
	return __viaddmin_u32(a, b, c);
}

// ---------------------------------------------------------------------------------------------
// DB re-layout: .edx clump (vector-major, byte = lane) -> chunked lane-major nibbles.
// Device layout: clump c = nchunks(c) * 16 uint4; piece (chunk, lane) holds the lane's codes for
// columns 32*chunk .. 32*chunk+31, column x in nibble x (byte x/2, low nibble = even x) -- the
// same nibble order as the file, so the transform is a pure byte transpose.
// ---------------------------------------------------------------------------------------------
__global__ void k_relayout(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
		const uint64_t *__restrict__ out_off, const uint32_t *__restrict__ clump_len,
		uint4 *__restrict__ out, uint32_t first, uint64_t in_base, uint32_t *__restrict__ meta_words) {
	uint32_t c = first + blockIdx.x;
	uint32_t flags = 0;
	uint32_t L = clump_len[c], nvec = (L + 1) >> 1, npieces = ((L + 31) >> 5) * 16;
	const uint8_t *src = in + (in_off[c] - in_base);
	uint4 *dst = out + out_off[c];
	for (uint32_t p = threadIdx.x; p < npieces; p += blockDim.x) {
		uint32_t chunk = p >> 4, lane = p & 15;
		uint32_t w[4] = {0, 0, 0, 0};
		#pragma unroll
		for (int j = 0; j < 16; ++j) {
			uint32_t v = chunk * 16 + j;
			uint32_t b = v < nvec ? src[(size_t)v * 16 + lane] : 0;
			w[j >> 2] |= b << (8 * (j & 3));
		}
		dst[p] = make_uint4(w[0], w[1], w[2], w[3]);
		#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const uint32_t ge6 = (((w[j] & 0x77777777u) + 0x22222222u) | w[j]) & 0x88888888u, ge5 = (((w[j] & 0x77777777u) + 0x33333333u) | w[j]) & 0x88888888u;
			if (ge6) flags |= 1u;
			if (ge5 ^ ge6) flags |= 2u;                                     // a nibble that is >= 5 but not >= 6
		}
	}
	if (flags) atomicOr(&meta_words[(size_t)c * 4 + 3], flags);            // ClumpMeta.flags (k_seed verifies words holding such codes by table)
}

// word `wi` (8 columns) of one lane: lanew points at the lane's first piece
__device__ __forceinline__ uint32_t lane_word(const uint32_t *lanew, uint32_t wi) {
	return __ldg(lanew + (size_t)(wi >> 2) * 64 + (wi & 3));
}

// ---------------------------------------------------------------------------------------------
// Query prep.
// k_qinfo:  raw (offset, budget, slot) arrays -> QInfo, validation, histogram of stretch lengths
//           hist[min(plen, 31)] over queries with k+1 <= SEED_NP_MAX, plen = len / (k+1).
// k_qprep:  nibble-packed copy of every query (two zero pad words, then 8 bases per word, base i in
//           nibble i & 7 -- the same nibble order as the DB) at word (off >> 3) + 3 * q, and the class:
//           cls = 1 (k_seed) iff the seed filter is on, every stretch is long enough for the batch's
//           window layout and the windows' bases are all plain A/C/G/T; else 0 (k_filter).
// ---------------------------------------------------------------------------------------------
#define SEED_NP_MAX 32                 // stretches per query the seed filter takes: 64 / stride windows per thread in the hash cache
// stride: reference windows are probed every `stride` columns (8 = word ends, 4 = also half words, 0 = off);
// w: window length in bases (8..16); hm: mask of the older word's nibbles inside the window;
// words: Bloom filter words per warp (power of two), shw = 32 - log2(words);
// amb_add: 0x2222.. flags reference codes >= 6, 0x3333.. codes >= 5 (N matches for free, -y) as "verify by table".
struct SeedLayout { uint32_t stride, w, hm, words, shw, amb_add, np_max; };   // np_max: most stretches per query the tag cache holds

__global__ void k_qinfo(const uint64_t *__restrict__ off, const uint16_t *__restrict__ budget, const uint32_t *__restrict__ slot,
		uint32_t nq, uint32_t nslots, QInfo *__restrict__ qi, uint32_t *__restrict__ hist, uint32_t *__restrict__ counters, uint64_t base_off) {
	__shared__ uint32_t sh[32];
	if (threadIdx.x < 32) sh[threadIdx.x] = 0;
	__syncthreads();
	uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q < nq) {
	uint64_t o = off[q], len = off[q + 1] - o;
	uint32_t k = budget[q], s = slot[q];
	if (off[q + 1] <= o || o < base_off || len > 0x7FFFFFFFull || k > 254 || s >= nslots) { atomicExch(&counters[C_ERR], q + 1); len = 1; k = 0; s = 0; o = base_off; }
	QInfo Q; Q.off = o - base_off; Q.len = (uint32_t)len; Q.slot = s; Q.k = (uint16_t)k; Q.P = filt_code(len, k); Q.cls = 0;
	qi[q] = Q;
	if (hist && k + 1 <= SEED_NP_MAX) atomicAdd(&sh[min((uint32_t)len / (k + 1), 31u)], 1u);
	}
	__syncthreads();
	if (hist && threadIdx.x < 32 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// BG_Q_PACKED4 input: nibble stream -> one code byte per base.  Thread i writes bases 16i .. 16i+15; `odd` = 1 when the
// first base of the range sits in the high nibble of packed[0].
__device__ __forceinline__ unsigned long long spread_nibbles(uint32_t w) {
	unsigned long long x = w;
	x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
	x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
	x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
	return x;
}
__global__ void k_unpack4(const uint8_t *__restrict__ packed, uint32_t odd, uint64_t nbases, uint8_t *__restrict__ out) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i * 16 >= nbases) return;
	const uint2 v = *(const uint2 *)(packed + i * 8);                 // the buffer is padded: reads past the last base stay inside it
	unsigned long long w = (unsigned long long)v.x | ((unsigned long long)v.y << 32);
	if (odd) w = (w >> 4) | ((unsigned long long)packed[i * 8 + 8] << 60);
	const unsigned long long lo = spread_nibbles((uint32_t)w), hi = spread_nibbles((uint32_t)(w >> 32));
	*(uint4 *)(out + i * 16) = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
}

// eight code bytes (two aligned-shifted words) -> eight nibbles
__device__ __forceinline__ uint32_t pack_nibbles(uint32_t b0, uint32_t b1) {
	unsigned long long x = ((unsigned long long)b0 | ((unsigned long long)b1 << 32)) & 0x0F0F0F0F0F0F0F0Full;
	x = (x | (x >> 4)) & 0x00FF00FF00FF00FFull;
	x = (x | (x >> 8)) & 0x0000FFFF0000FFFFull;
	x = (x | (x >> 16));
	return (uint32_t)x;
}

// nibbles of w that are not a plain base (codes 1..4): bit 3 of the nibble set
__device__ __forceinline__ uint32_t nonplain_nibbles(uint32_t w) {
	const uint32_t a = w & 0x77777777u;
	return (~(a + 0x77777777u) | (a + 0x33333333u) | w) & 0x88888888u;
}

__global__ void k_qprep(const uint8_t *__restrict__ codes, QInfo *__restrict__ qi, uint32_t nq, SeedLayout SL,
		uint32_t *__restrict__ qnib, uint32_t *__restrict__ nseed, uint32_t *__restrict__ unseeded) {
	const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	bool seed = false; uint32_t nst = 0, notplain = 0;
	if (q < nq) {
		const QInfo Q = qi[q];
		// the code bytes through aligned 32-bit loads (`codes` is 256-byte aligned and padded), shifted into place
		const uint32_t *a = (const uint32_t *)(codes + (Q.off & ~3ull));
		const uint32_t sh = (uint32_t)(Q.off & 3) * 8;
		uint32_t *W = qnib + (Q.off >> 3) + 3ull * q;
		W[0] = 0; W[1] = 0;
		uint32_t w0 = a[0], w1 = a[1];
		for (uint32_t j = 0, k = 0; j < Q.len; j += 8, k += 2) {
			const uint32_t w2 = a[k + 2];
			const uint32_t b0 = __funnelshift_r(w0, w1, sh), b1 = __funnelshift_r(w1, w2, sh);
			w0 = w2; w1 = a[k + 3];
			uint32_t w = pack_nibbles(b0, b1);
			if (j + 8 > Q.len) { w &= (1u << (4 * (Q.len - j))) - 1u; notplain |= nonplain_nibbles(w | (0x11111111u << (4 * (Q.len - j)))); }
			else notplain |= nonplain_nibbles(w);
			W[2 + (j >> 3)] = w;
		}
		const uint32_t np = Q.k + 1u, plen = Q.len / np;
		seed = SL.stride && np <= SL.np_max && plen >= SL.w + SL.stride - 1;
		// every base the windows cover must be a plain A/C/G/T (codes 1..4): checked on the packed words just written
		const uint32_t span = SL.w + SL.stride - 1;
		for (uint32_t p = 0; p < np && seed; ++p) {
			const uint32_t E = (p + 1) * plen, B = E - span;
			for (uint32_t wi = B >> 3; wi <= (E - 1) >> 3; ++wi) {
				const uint32_t w = W[2 + wi];
				uint32_t bad = ((((w & 0x77777777u) + 0x33333333u) | w) | ((w - 0x11111111u) & ~w)) & 0x88888888u;   // nibble >= 5, or 0 (or just above a 0)
				const uint32_t lo = wi * 8 < B ? B - wi * 8 : 0u, hi = wi * 8 + 8 > E ? E - wi * 8 : 8u;            // nibbles [lo, hi) of this word are inside
				uint32_t keep = hi == 8 ? 0xFFFFFFFFu : (1u << (4 * hi)) - 1u;
				keep &= ~((1u << (4 * lo)) - 1u);
				// a zero nibble raises the flag of the nibbles above it as well; those are inside the region or beyond its end, where a
				// spurious flag could only come from a zero below, itself inside the word -- so test from the region start upward
				bad &= keep;
				if (bad) {                                                     // confirm nibble by nibble (rare)
					for (uint32_t i = lo; i < hi; ++i) { const uint32_t c = (w >> (4 * i)) & 15; if (c < 1 || c > 4) { seed = false; break; } }
					if (!seed) break;
				}
			}
		}
		qi[q].cls = (uint8_t)((seed ? 1u : 0u) | (notplain ? 0u : 2u));
		if (seed) nst = np;
	}
	const uint32_t m = __ballot_sync(0xFFFFFFFFu, seed), tot = __reduce_add_sync(0xFFFFFFFFu, nst), mx = __reduce_max_sync(0xFFFFFFFFu, nst);
	if (m && (threadIdx.x & 31) == 0) { atomicAdd(nseed, (uint32_t)__popc(m)); atomicAdd(nseed + 1, tot); atomicMax(nseed + 2, mx); }   // seeded queries, their stretches (sum, max)
	const uint32_t u = __ballot_sync(0xFFFFFFFFu, q < nq && !seed);
	if (unseeded && u && (threadIdx.x & 31) == 0) atomicAdd(unseeded, (uint32_t)__popc(u));
}

// ---------------------------------------------------------------------------------------------
// Diagonal clusters: a tiny sorted set of disjoint intervals (rare path, local memory is fine).
// ---------------------------------------------------------------------------------------------
#define CLUS_MAX 12                    // (the count travels in 4 bits of Surv::w_lane; 4 was too few on 14 kb lanes of sheared genomes: distant clusters fused into bands thousands of cells wide)
struct Clus { int lo[CLUS_MAX + 1], hi[CLUS_MAX + 1]; int n; };

__device__ __noinline__ void clus_add(Clus &C, int lo, int hi) {
	int tl[CLUS_MAX + 1], th[CLUS_MAX + 1], out = 0; bool placed = false;
	for (int i = 0; i < C.n; ++i) {
		if (C.hi[i] + 1 < lo) { tl[out] = C.lo[i]; th[out] = C.hi[i]; ++out; }
		else if (C.lo[i] > hi + 1) {
			if (!placed) { tl[out] = lo; th[out] = hi; ++out; placed = true; }
			tl[out] = C.lo[i]; th[out] = C.hi[i]; ++out;
		} else { lo = min(lo, C.lo[i]); hi = max(hi, C.hi[i]); }       // overlapping or adjacent: absorb
	}
	if (!placed) { tl[out] = lo; th[out] = hi; ++out; }
	if (out > CLUS_MAX) {                                              // too many: fuse the two closest neighbours
		int bi = 0, bg = INT32_MAX;
		for (int i = 0; i + 1 < out; ++i) { int g = tl[i + 1] - th[i]; if (g < bg) { bg = g; bi = i; } }
		th[bi] = th[bi + 1];
		for (int i = bi + 1; i + 1 < out; ++i) { tl[i] = tl[i + 1]; th[i] = th[i + 1]; }
		--out;
	}
	for (int i = 0; i < out; ++i) { C.lo[i] = tl[i]; C.hi[i] = th[i]; }
	C.n = out;
}

__device__ __noinline__ void emit_clusters(const Clus &C, uint32_t task, uint32_t lane, Surv *surv, uint32_t surv_cap, uint32_t *counters) {
	if (!C.n) return;
	const uint32_t base = atomicAdd(&counters[C_SURV], (uint32_t)C.n);
	for (int s = 0; s < C.n; ++s) {
		const uint32_t W = (uint32_t)(C.hi[s] - C.lo[s] + 1);
		uint32_t scratch = 0;
		if (W > 64) atomicMax(&counters[C_SCRATCH], W);
		if (base + s < surv_cap) {
			Surv v; v.task = task; v.lo = C.lo[s]; v.w_lane = (W << 8) | ((s == 0 ? (uint32_t)C.n : 0u) << 4) | lane; v.scratch = scratch;
			surv[base + s] = v;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Phase A1: pigeonhole seed filter.  128 threads = 4 warps, each warp owns a chunk of consecutive runs.
//
// Why a probe every `stride` columns is enough.  Let a query of length m with budget k be cut into k+1
// stretches of plen = m / (k+1) bases (stretch p ends at query offset E = (p+1) * plen).  An alignment
// with <= k errors leaves at least one stretch without error: its plen bases sit on ONE diagonal opposite
// plen consecutive reference columns [a, a + plen).  With plen >= w + stride - 1 those columns contain a
// window of w columns that ends on a multiple of `stride`; the query bases opposite it are a window of w
// bases ending at E - j for some 0 <= j < stride.  So: put the `stride` windows ending at E, E-1, ..,
// E-stride+1 of every stretch into a set, probe the set with the reference window ending at every
// multiple of `stride`, and every alignment within budget is found, with seed diagonal
// d = (reference end column) - (E - j); the whole alignment then lies within diagonals [d-k, d+k].
// The set is a blocked Bloom filter (one 32-bit word, two bits) per warp in shared memory, keyed by the
// window's nibbles; a set bit is only a candidate, the verification below compares the nibbles.
// Reference codes that can match a plain base without being equal to it (IUPAC codes, N under -y)
// bypass the filter: their words are always flagged and verified through the scoring table.
// ---------------------------------------------------------------------------------------------
// One record per clump for k_seed: a single 16-byte load instead of two dependent ones.
struct ClumpMeta { uint64_t off; uint32_t len; uint32_t flags; };   // off in uint4 units; flags: 1 = some code >= 6, 2 = some code == 5
struct SeedArgs {
	const uint4 *db; const ClumpMeta *meta;
	const QInfo *qi; const uint32_t *qnib; Work W; SeedLayout SL;
	uint64_t nwork; uint32_t chunk;                 // run indices to enumerate, runs per warp
	uint32_t npmax, hb, stage;                      // stretches per query the window table holds, its hash buckets; bytes per staging buffer
	Surv *surv; uint32_t surv_cap; uint32_t *counters;
	uint32_t m16[8];                                // match sets: bit r of half-word q = (S[q][r] == 0)
};

// work index -> run.  Explicit lists are walked as given (a bunch's clump visits are consecutive and share
// their queries); the implicit all-vs-all list is walked tile-major for the same reason, while the run id
// (and with it the task id the host sees) stays clump-major.
__device__ __forceinline__ bool get_work(const Work &W, uint64_t i, uint64_t &r, uint32_t &c, uint32_t &q0, uint32_t &n) {
	if (W.runs) r = i;
	else { const uint64_t tile = i / W.num_clumps, cl = i - tile * W.num_clumps; r = cl * W.ntiles + tile; }
	return get_run(W, r, c, q0, n);
}

#define HASH_C1 0x9E3779B1u
#define HASH_C2 0x85EBCA77u
__device__ __forceinline__ uint32_t seed_hash(uint32_t newer, uint32_t older_masked) { return newer * HASH_C1 + older_masked * HASH_C2; }
// blocked Bloom filter: word = top bits of the hash, two bit positions from its low ten bits
__device__ __forceinline__ uint32_t bloom_bits(uint32_t h) { return (1u << (h & 31)) | (1u << ((h >> 5) & 31)); }
__device__ __forceinline__ uint32_t bloom_test(uint32_t word, uint32_t h) { return (word >> (h & 31)) & (word >> ((h >> 5) & 31)) & 1u; }
__device__ __forceinline__ uint32_t amb_nibbles(uint32_t w, uint32_t add) { return (((w & 0x77777777u) + add) | w) & 0x88888888u; }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ uint4 lds128(uint32_t addr) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v; }

// mbarrier + bulk copy (TMA, 1-D): one elected thread arms the barrier with the byte count and issues the copy
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	asm volatile("{\n\t.reg .pred p;\n\tMB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra MB_DONE;\n\tbra MB_WAIT;\n\tMB_DONE:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}

// Window (p, j) of a query: stretch p ends at offset E = (p+1) * plen; the window's 16 bases end at E - j.
// Returns the newer 8 bases, the older 8 (unmasked) and the end offset.
struct QWin { uint32_t kn, ko, y1; };
struct QStretch { uint32_t r0, r1, r2, E; };
__device__ __forceinline__ QStretch stretch_of(const uint32_t *__restrict__ Wq, uint32_t plen, uint32_t p) {
	QStretch S; S.E = (p + 1) * plen;
	const uint32_t wI = (S.E + 8) >> 3, sh = ((S.E + 8) & 7) * 4;          // bases [E-8, E) start at padded nibble E+8
	const uint32_t a = __ldg(Wq + wI - 2), b = __ldg(Wq + wI - 1), c = __ldg(Wq + wI), d = __ldg(Wq + wI + 1);
	S.r0 = __funnelshift_r(a, b, sh); S.r1 = __funnelshift_r(b, c, sh); S.r2 = __funnelshift_r(c, d, sh);
	return S;
}
__device__ __forceinline__ QWin window_of(const QStretch &S, uint32_t j) {
	QWin w; w.kn = __funnelshift_l(S.r1, S.r2, 4 * j); w.ko = __funnelshift_l(S.r0, S.r1, 4 * j); w.y1 = S.E - j; return w;
}

// nibble-by-nibble comparison through the match sets (windows holding ambiguous reference codes only)
__device__ __noinline__ bool window_matches_table(const uint32_t *sM, uint32_t kn, uint32_t ko, uint32_t rn, uint32_t ro, uint32_t w) {
	for (uint32_t t = 16 - w; t < 16; ++t) {
		const uint32_t qn = ((t < 8 ? ko >> (4 * t) : kn >> (4 * (t - 8)))) & 15, rr = ((t < 8 ? ro >> (4 * t) : rn >> (4 * (t - 8)))) & 15;
		if (!((sM[qn] >> rr) & 1)) return false;
	}
	return true;
}

// Shared memory of one block, in 32-bit words (every part a multiple of 4 words).  A block walks its runs in rounds
// of G (= groups per block, 16 threads each, one run per group); the runs of a round that belong to one bunch share
// ONE window table, built by the whole block:
//   sM     [16]             match sets of the scoring table
//   bits   [words]          Bloom filter over the windows of the bunch's queries
//   head   [hb]             hash buckets of the window table: entry index + 1, 0 = empty
//   tag    [ne]             full 32-bit hash of window e = (query * npmax + stretch) * stride + j
//   next   [ne / 2]         chain links (16 bit)
//   str    [4 * 16 * npmax] the stretch registers (r0, r1, r2, E) of (query, stretch); E = 0: no such stretch
//   kq     [16]             budgets of the bunch's queries
//   keys   [16]             (query0, nq) of the run each group holds in this round
//   per group: stage [2 * stage / 4] two staging buffers for clumps (bulk copies), mbar [4] two mbarriers
struct SeedSmem { uint32_t bits, head, tag, next, str, kq, keys, groups, gwords, total; };
__host__ __device__ __forceinline__ SeedSmem seed_smem(uint32_t words, uint32_t hb, uint32_t npmax, uint32_t stride, uint32_t stage, uint32_t G) {
	SeedSmem M; const uint32_t ne = 16 * npmax * stride;
	M.bits = 16; M.head = M.bits + words; M.tag = M.head + hb; M.next = M.tag + ne; M.str = M.next + ((ne / 2 + 3) & ~3u);
	M.kq = M.str + 64 * npmax; M.keys = M.kq + 16; M.groups = M.keys + 16; M.gwords = 2 * (stage / 4) + 4; M.total = M.groups + G * M.gwords;
	return M;
}

// A thread's seeds of one run (it owns one reference lane): a short list, and per query the hull of its
// seed diagonals for the (rare) case that the list overflows.
#define SEED_LIST 24
struct LaneSeeds { uint32_t q[SEED_LIST]; int d[SEED_LIST]; int n; int hlo[16], hhi[16]; };

__device__ __noinline__ void lane_seeds_overflow(LaneSeeds &L, uint32_t qi, int dg) {
	if (L.n == SEED_LIST) {                                                // first overflow: hulls of what the list holds
		for (int i = 0; i < 16; ++i) { L.hlo[i] = INT32_MAX; L.hhi[i] = INT32_MIN; }
		for (int i = 0; i < SEED_LIST; ++i) { L.hlo[L.q[i]] = min(L.hlo[L.q[i]], L.d[i]); L.hhi[L.q[i]] = max(L.hhi[L.q[i]], L.d[i]); }
		L.n = SEED_LIST + 1;
	}
	L.hlo[qi] = min(L.hlo[qi], dg); L.hhi[qi] = max(L.hhi[qi], dg);
}

// clusters of one (query, lane) from the list -> survivors
__device__ __noinline__ void lane_seeds_emit(LaneSeeds &L, const uint32_t *kq, uint32_t task0, uint32_t lane, Surv *surv, uint32_t surv_cap, uint32_t *counters) {
	if (L.n > SEED_LIST) {
		for (uint32_t qi = 0; qi < 16; ++qi) if (L.hlo[qi] <= L.hhi[qi]) {
			Clus C; C.n = 1; C.lo[0] = L.hlo[qi] - (int)kq[qi]; C.hi[0] = L.hhi[qi] + (int)kq[qi];
			emit_clusters(C, task0 + qi, lane, surv, surv_cap, counters);
		}
		return;
	}
	uint32_t done = 0;
	for (int i = 0; i < L.n; ++i) {
		const uint32_t qi = L.q[i];
		if (done >> qi & 1) continue;
		done |= 1u << qi;
		const int k = (int)kq[qi];
		Clus C; C.n = 0;
		for (int s = 0; s <= CLUS_MAX; ++s) { C.lo[s] = 0; C.hi[s] = 0; }
		for (int j = i; j < L.n; ++j) if (L.q[j] == qi) clus_add(C, L.d[j] - k, L.d[j] + k);
		emit_clusters(C, task0 + qi, lane, surv, surv_cap, counters);
	}
}

// Block-wide barrier for k_seed.  The two 16-thread groups of a warp take different paths between barriers (one scans and
// emits, the other may not), and bar.sync is an ALIGNED barrier: every thread of a warp has to execute it together.  Leaving
// the reconvergence to the compiler made the kernel fault or lose seeds depending on unrelated code changes
// (profiles/r1d_known_issue.txt); the explicit __syncwarp() makes the warp whole again before it arrives.
__device__ __forceinline__ void block_barrier() { __syncwarp(); __syncthreads(); }
__device__ __forceinline__ bool block_barrier_or(bool p) { __syncwarp(); return __syncthreads_or(p) != 0; }

template <int STRIDE, bool FULLW>   // FULLW: 16-base windows (no mask on the older word)
__global__ void __launch_bounds__(128, 8) k_seed(SeedArgs A) {
	extern __shared__ __align__(128) uint32_t smem[];
	const uint32_t G = blockDim.x >> 4;                                    // groups of the block = runs per round
	const uint32_t grp = threadIdx.x >> 4, l = threadIdx.x & 15;          // group, reference lane within the run
	const uint32_t gmask = 0xFFFFu << (threadIdx.x & 16);                  // the group's threads within their warp
	const SeedSmem M = seed_smem(A.SL.words, A.hb, A.npmax, STRIDE, A.stage, G);
	uint32_t *sM = smem, *bits = smem + M.bits, *head = smem + M.head, *tag = smem + M.tag, *str = smem + M.str, *kq = smem + M.kq;
	uint16_t *nxt = (uint16_t *)(smem + M.next);
	unsigned long long *keys = (unsigned long long *)(smem + M.keys);
	uint32_t *gbase = smem + M.groups + grp * M.gwords;
	const uint32_t bits_s = (uint32_t)__cvta_generic_to_shared(bits), stg_s = (uint32_t)__cvta_generic_to_shared(gbase);
	const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(gbase + 2 * (A.stage / 4));
	if (threadIdx.x < 16) sM[threadIdx.x] = (A.m16[threadIdx.x >> 1] >> (16 * (threadIdx.x & 1))) & 0xFFFFu;
	if (l == 0) { mbar_init(bar_s, 1); mbar_init(bar_s + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
	__syncthreads();
	const uint64_t b0 = (uint64_t)blockIdx.x * G * A.chunk;               // the block's runs: b0 .. b0 + G * chunk
	const uint32_t HM = FULLW ? 0xFFFFFFFFu : A.SL.hm, SHW = A.SL.shw, ADD = A.SL.amb_add, ambsel = ADD == 0x33333333u ? 3u : 1u;
	const uint32_t HBM = A.hb - 1, NPM = A.npmax;
	const unsigned long long NOKEY = ~0ull;
	unsigned long long tableK = NOKEY;                                     // the bunch the shared window table was built for (block-uniform)
	bool table_any = false;                                               // ... and whether any of its queries is seeded

	struct Desc { uint64_t r; uint32_t c, q0, n; bool ok; ClumpMeta M; };
	auto load_desc = [&](uint64_t i, Desc &D) {
		D.ok = i < A.nwork && get_work(A.W, i, D.r, D.c, D.q0, D.n);
		if (D.ok) { const uint4 m = __ldg((const uint4 *)(A.meta + D.c)); D.M.off = (uint64_t)m.x | ((uint64_t)m.y << 32); D.M.len = m.z; D.M.flags = m.w; }
	};
	// bulk copy of a clump into staging buffer b (the whole group calls; its first thread issues)
	auto stage_issue = [&](const Desc &D, uint32_t b) -> bool {
		const uint32_t bytes = ((D.M.len + 31) >> 5) * 256;
		if (!D.ok || bytes > A.stage) return false;
		__syncwarp(gmask);                                                 // everyone is done reading buffer b
		if (l == 0) { mbar_expect_tx(bar_s + 8 * b, bytes); bulk_g2s(stg_s + b * A.stage, A.db + D.M.off, bytes, bar_s + 8 * b); }
		return true;
	};
	Desc cur, nxtd;
	load_desc(b0 + grp, cur);
	bool cur_staged = stage_issue(cur, 0);
	uint32_t buf = 0, phase = 0;

	for (uint32_t round = 0; round < A.chunk; ++round, cur = nxtd, buf ^= 1) {
		if (b0 + (uint64_t)round * G >= A.nwork) break;                    // block-uniform
		bool nxt_staged = false; nxtd.ok = false;
		if (round + 1 < A.chunk) { load_desc(b0 + (uint64_t)(round + 1) * G + grp, nxtd); nxt_staged = stage_issue(nxtd, buf ^ 1); }
		const bool staged = cur_staged; cur_staged = nxt_staged;
		const unsigned long long mykey = cur.ok ? ((unsigned long long)cur.q0 << 8) | cur.n : NOKEY;
		if (l == 0) keys[grp] = mykey;
		block_barrier();                                                   // keys published; the previous round is done with the table
		uint32_t processed = 0;
		for (uint32_t g = 0; g < G; ++g) if (keys[g] == NOKEY) processed |= 1u << g;
		if (processed == (1u << G) - 1) block_barrier();                   // nothing to do this round: keep the next round's keys behind a barrier all the same
		while (processed != (1u << G) - 1) {
			// ---- the first pending bunch of the round: every run of it is scanned against one table ----
			const unsigned long long K = keys[__ffs(~processed) - 1];
			for (uint32_t g = 0; g < G; ++g) if (keys[g] == K) processed |= 1u << g;
			const bool active = mykey == K;
			const uint32_t q0 = (uint32_t)(K >> 8), n = (uint32_t)(K & 255);
			if (K != tableK) {                                             // (a bunch with more runs than a round keeps its table)
				// build: thread = (query t & 15, window phase t >> 4)
				for (uint32_t w = threadIdx.x * 4; w < A.SL.words + A.hb; w += blockDim.x * 4) *(uint4 *)(bits + w) = make_uint4(0, 0, 0, 0);   // bits and head are adjacent
				const uint32_t qi = threadIdx.x & 15, sub = threadIdx.x >> 4;
				bool act = false; QInfo Q; Q.len = 0; Q.k = 0; Q.off = 0;
				if (qi < n) { Q = A.qi[q0 + qi]; act = (Q.cls & 1) != 0; }
				if (sub == 0) kq[qi] = Q.k;
				block_barrier();
				{
					const uint32_t np = act ? Q.k + 1u : 0u, plen = act ? Q.len / np : 0u;
					const uint32_t *Wq = A.qnib + (Q.off >> 3) + 3ull * (q0 + qi);
					for (uint32_t wn = sub; wn < NPM * STRIDE; wn += G) {
						const uint32_t p = wn / STRIDE, j = wn % STRIDE;
						if (p < np) {
							const QStretch S = stretch_of(Wq, plen, p);
							const QWin w = window_of(S, j);
							const uint32_t hv = seed_hash(w.kn, w.ko & HM), e = (qi * NPM + p) * STRIDE + j;
							atomicOr(&bits[hv >> SHW], bloom_bits(hv));
							tag[e] = hv;
							nxt[e] = (uint16_t)atomicExch(&head[(hv >> 10) & HBM], e + 1);
							if (j == 0) *(uint4 *)(str + (qi * NPM + p) * 4) = make_uint4(S.r0, S.r1, S.r2, S.E);
						} else if (j == 0) *(uint4 *)(str + (qi * NPM + p) * 4) = make_uint4(0, 0, 0, 0);
					}
				}
				table_any = block_barrier_or(act);
				tableK = K;
			}
			const bool anyact = table_any;
			if (active) {
				if (staged) { mbar_wait(bar_s + 8 * buf, (phase >> buf) & 1u); phase ^= 1u << buf; }
				if (anyact) {
		// ---- scan: this thread streams its lane, one probe per `STRIDE` columns ----
		const uint32_t L = cur.M.len, nchunks = (L + 31) >> 5;
		const uint32_t gs = nchunks <= 8 ? 0u : max(2u, 32u - __clz((nchunks * 4 - 1) >> 5));   // 2^gs words per mask bit
		const bool amb_on = (cur.M.flags & ambsel) != 0;
		const uint4 *gp = A.db + cur.M.off;                                // piece (chunk, lane) at gp[chunk * 16 + lane]
		const uint32_t sp = stg_s + buf * A.stage;
		auto word_at = [&](uint32_t wi) -> uint32_t {
			return staged ? lds32(sp + ((wi >> 2) * 16 + l) * 16 + (wi & 3) * 4) : __ldg((const uint32_t *)(gp + (size_t)(wi >> 2) * 16 + l) + (wi & 3));
		};
		uint32_t mask = 0;
		auto scan = [&](auto staged_c, auto amb_c) {
			constexpr bool ST = decltype(staged_c)::value, AMB = decltype(amb_c)::value;
			auto ld = [&](uint32_t ck) -> uint4 { return ST ? lds128(sp + (ck * 16 + l) * 16) : __ldg(gp + (size_t)ck * 16 + l); };
			uint32_t prev = 0, prev2 = 0, ambp = 0, ambp2 = 0;
			uint4 pc = ld(0), pn = nchunks > 1 ? ld(1) : make_uint4(0, 0, 0, 0);   // two chunks (64 columns) in flight
			for (uint32_t ck = 0; ck < nchunks; ++ck) {
				const uint4 cw = pc;
				pc = pn;
				if (ck + 2 < nchunks) pn = ld(ck + 2);
				const uint32_t ws[4] = {cw.x, cw.y, cw.z, cw.w};
				uint32_t m4 = 0;
				#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const uint32_t cu = ws[j];
					const uint32_t h8 = seed_hash(cu, prev & HM);
					uint32_t hit = bloom_test(lds32(bits_s + ((h8 >> SHW) << 2)), h8);
					if (STRIDE == 4) {
						const uint32_t h4 = seed_hash(__funnelshift_r(prev, cu, 16), __funnelshift_r(prev2, prev, 16) & HM);
						hit |= bloom_test(lds32(bits_s + ((h4 >> SHW) << 2)), h4);
					}
					if (AMB) {
						const uint32_t ambc = amb_nibbles(cu, ADD);
						if (ambc | ambp | (STRIDE == 4 ? ambp2 : 0u)) hit = 1;
						ambp2 = ambp; ambp = ambc;
					}
					m4 |= hit << j;
					prev2 = prev; prev = cu;
				}
				const uint32_t rel = ck * 4;
				mask |= (gs ? (uint32_t)(m4 != 0) : m4) << (rel >> gs);
			}
		};
		if (staged) { if (amb_on) scan(std::true_type{}, std::true_type{}); else scan(std::true_type{}, std::false_type{}); }
		else        { if (amb_on) scan(std::false_type{}, std::true_type{}); else scan(std::false_type{}, std::false_type{}); }

		// ---- verify this lane's flagged words against the window table; seeds -> clusters -> survivors ----
		if (mask) {
			LaneSeeds LS; LS.n = 0;
			auto seed = [&](uint32_t sq, int dg) {
				if (LS.n < SEED_LIST) { LS.q[LS.n] = sq; LS.d[LS.n] = dg; ++LS.n; }
				else lane_seeds_overflow(LS, sq, dg);
			};
			const uint32_t nw = nchunks * 4;
			while (mask) {
				const uint32_t s = __ffs(mask) - 1; mask &= mask - 1;
				for (uint32_t wi = s << gs; wi < min(nw, (s + 1) << gs); ++wi) {
					const uint32_t cu = word_at(wi), pv = wi >= 1 ? word_at(wi - 1) : 0u, pv2 = (STRIDE == 4 && wi >= 2) ? word_at(wi - 2) : 0u;
					#pragma unroll
					for (int e = STRIDE; e <= 8; e += STRIDE) {
						const uint32_t rn = e == 8 ? cu : __funnelshift_r(pv, cu, 16), ro = (e == 8 ? pv : __funnelshift_r(pv2, pv, 16)) & HM;
						const int x1 = (int)(wi * 8 + e);
						if (amb_on && (amb_nibbles(rn, ADD) | amb_nibbles(ro, ADD))) {       // IUPAC codes in the window: every window of the bunch, through the table
							for (uint32_t si = 0; si < 16 * NPM; ++si) {
								const uint4 rec = *(const uint4 *)(str + si * 4);
								if (!rec.w) continue;
								QStretch S; S.r0 = rec.x; S.r1 = rec.y; S.r2 = rec.z; S.E = rec.w;
								for (uint32_t j = 0; j < (uint32_t)STRIDE; ++j) {
									const QWin w = window_of(S, j);
									if (window_matches_table(sM, w.kn, w.ko & HM, rn, ro, A.SL.w)) seed(si / NPM, x1 - (int)w.y1);
								}
							}
						} else {
							const uint32_t hv = seed_hash(rn, ro);
							for (uint32_t en = head[(hv >> 10) & HBM]; en; en = nxt[en - 1]) {
								if (tag[en - 1] != hv) continue;
								const uint32_t si = (en - 1) / STRIDE, j = (en - 1) % STRIDE;
								const uint4 rec = *(const uint4 *)(str + si * 4);
								QStretch S; S.r0 = rec.x; S.r1 = rec.y; S.r2 = rec.z; S.E = rec.w;
								const QWin w = window_of(S, j);
								if (w.kn == rn && (w.ko & HM) == ro) seed(si / NPM, x1 - (int)w.y1);
							}
						}
					}
				}
			}
			if (LS.n) {
				const uint32_t task0 = (uint32_t)((cur.r + A.W.run_base) * BG_RUN_MAX);
				bool simple = LS.n <= SEED_LIST;                               // usual case: one query, one cluster
				int dlo = LS.d[0], dhi = LS.d[0];
				if (simple) for (int z = 1; z < LS.n; ++z) { simple = simple && LS.q[z] == LS.q[0]; dlo = min(dlo, LS.d[z]); dhi = max(dhi, LS.d[z]); }
				const int k0 = (int)kq[LS.q[0] & 15];
				if (simple && dhi - dlo <= 2 * k0 + 1) {
					const uint32_t W = (uint32_t)(dhi - dlo + 2 * k0 + 1);
					const uint32_t slot = atomicAdd(&A.counters[C_SURV], 1u);
					uint32_t scratch = 0;
					if (W > 64) atomicMax(&A.counters[C_SCRATCH], W);
					if (slot < A.surv_cap) { Surv v; v.task = task0 + LS.q[0]; v.lo = dlo - k0; v.w_lane = (W << 8) | (1u << 4) | l; v.scratch = scratch; A.surv[slot] = v; }
				} else lane_seeds_emit(LS, kq, task0, l, A.surv, A.surv_cap, A.counters);
			}
		}
				}
			}
			block_barrier();                                               // the table is free for the next bunch of the round
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Phase A1, warp form (the default).  The same pigeonhole filter as k_seed above, reorganised so that
// nothing in the steady state waits on a block-wide barrier:
//   * a WARP owns a bunch (the <= 16 queries that share a candidate list, burst.c:4077-4157): its window
//     set lives in the warp's private slice of shared memory -- a blocked two-bit Bloom filter (the scan
//     filter) and chained buckets of window ids (the exact verification; one atomicExch per insertion)
//     -- built by the 32 lanes together and kept for every clump visit (run) of the bunch; only
//     __syncwarp() orders it;
//   * the two half-warps scan two runs of the bunch at a time, a thread per reference lane.  Clumps come
//     through TMA: the first lane of a half-warp issues ONE bulk copy (cp.async.bulk, <= NCH * 256 bytes)
//     of the next item into the half-warp's second staging buffer and arms its mbarrier, then everyone
//     probes the current buffer (128-bit conflict-free shared loads); the exact verification of a hit
//     re-reads its words from the same buffer, never from global memory.  The run records + clump records
//     of a whole segment (<= 32 runs) are fetched by one coalesced load per lane before the first is needed;
//   * a probe is ten instructions: two multiply-adds (hash of the 16-base window), shift + address,
//     one shared load, three shifts + an AND that test the two bits, one funnel shift that appends the
//     result to the hit mask;
//   * survivors of the usual kind (one query, one diagonal cluster) leave through one atomicAdd per warp.
// Work is cut into chunks of `chunk` consecutive runs, chunks are dealt to warps round-robin; a bunch is
// scanned by the warp whose chunk holds its first run (bunches longer than SEEDW_SPLIT runs are split), so
// a bunch's table is built once.
// ---------------------------------------------------------------------------------------------
#define SEEDW_WARPS 4
#define SEEDW_SPLIT 256
struct SeedWArgs {
	const uint4 *db; const ClumpMeta *meta;
	const QInfo *qi; const uint32_t *qnib; Work W; SeedLayout SL;
	uint64_t nwork; uint32_t chunk;                 // run indices to enumerate, runs per chunk
	uint32_t npmax, lbits, hslots;                  // stretches per query in the table; log2 of the bitmap bits; table slots (power of two)
	Surv *surv; uint32_t surv_cap; uint32_t *counters;
	uint32_t m16[8];
};
// per-warp shared memory in words: filter | bucket heads | chain links (16 bit) | stretch records | budgets | 4 staging buffers (2 per half-warp) |
// 4 mbarriers | 32 survivors waiting to leave | 64 flagged words waiting for verification | per-lane seed records | queue fill.  hslots = buckets (power of two), ne = windows the table can hold = 16 * npmax * stride
__host__ __device__ __forceinline__ uint32_t seedw_warp_words(uint32_t lbits, uint32_t hslots, uint32_t npmax, uint32_t stride, uint32_t nch) {
	return (1u << (lbits - 5)) + hslots + ((8 * npmax * stride + 3) & ~3u) + 64 * npmax + 16 + 4 * nch * 64 + 8 + 128 + 64 + 128 + 4;
}

template <int STRIDE, bool FULLW, int NCH, int FB, int VM>   // FB: bits per window in the filter (1 or 2); VM: verification of flagged words, 0 = queue + helper lanes + shared-memory atomics, 1 = one pass per flagged lane, results by warp reduction
__global__ void __launch_bounds__(SEEDW_WARPS * 32, 4) k_seedw(SeedWArgs A) {
	extern __shared__ __align__(16) uint32_t smem[];
	constexpr uint32_t FULL = 0xFFFFFFFFu;
	constexpr int NW = NCH * 4;                                           // words (of 8 columns) per item
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, l = lane & 15;
	const uint32_t BW = 1u << (A.lbits - 5), HSM = A.hslots - 1, NPM = A.npmax;
	const uint32_t NPI = 65536u / NPM + 1u;                                  // si / NPM == (si * NPI) >> 16 for si < 16 * NPM, NPM <= 32 (a runtime division costs ~20 instructions per match)
	uint32_t *sM = smem;
	uint32_t *bits = smem + 16 + warp * seedw_warp_words(A.lbits, A.hslots, NPM, STRIDE, NCH), *slots = bits + BW;   // slots: bucket heads, entry + 1 (0 = empty)
	uint16_t *nxt = (uint16_t *)(slots + A.hslots);                        // chain links
	uint32_t *str = slots + A.hslots + ((8 * NPM * STRIDE + 3) & ~3u), *kq = str + 64 * NPM;
	constexpr uint32_t ITEM = NCH * 256;                                   // bytes of one staging buffer
	uint32_t *stage0 = kq + 16;                                            // this warp's staging: half h uses buffers 2h, 2h+1
	const uint32_t bits_s = (uint32_t)__cvta_generic_to_shared(bits);
	const uint32_t stg_s = (uint32_t)__cvta_generic_to_shared(stage0) + half * 2 * ITEM;
	const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(stage0 + 4 * NCH * 64) + half * 16;
	if (threadIdx.x < 16) sM[threadIdx.x] = (A.m16[threadIdx.x >> 1] >> (16 * (threadIdx.x & 1))) & 0xFFFFu;
	if (l == 0) { mbar_init(bar_s, 1); mbar_init(bar_s + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
	__syncthreads();
	uint32_t cb = 0, phase = 0;                                            // staging buffer that holds the current item (warp-uniform); per-buffer mbarrier parity
	uint4 *sbuf = (uint4 *)(stage0 + 4 * NCH * 64 + 8); uint32_t scount = 0;   // survivors collected by the warp; they leave 32 at a time through one atomicAdd
	uint32_t *hq = stage0 + 4 * NCH * 64 + 8 + 128;                        // flagged words waiting for verification: owner lane << 8 | half-word window << 7 | word of the item
	constexpr uint32_t HQ = 64;
	uint32_t *ost = hq + HQ, *hqn = ost + 128;                             // per lane: {query, lowest, highest seed diagonal, conflict flag} of its current run; queue fill
	constexpr uint32_t NOQ = 0xFFFFFFFFu;
	const uint32_t stgw_s = (uint32_t)__cvta_generic_to_shared(stage0);   // staging of the whole warp (a helper lane reads the owner's buffer)
	auto flush = [&]() {
		uint32_t base = 0;
		if (lane == 0) base = atomicAdd(&A.counters[C_SURV], scount);
		base = __shfl_sync(FULL, base, 0);
		if (lane < scount && base + lane < A.surv_cap) ((uint4 *)A.surv)[base + lane] = sbuf[lane];
		scount = 0;
		__syncwarp();
	};
	const uint32_t HM = FULLW ? FULL : A.SL.hm, SHB = 32 - (A.lbits - 5), ADD = A.SL.amb_add, ambsel = ADD == 0x33333333u ? 3u : 1u;
	const unsigned long long NOKEY = ~0ull;
	unsigned long long tableK = NOKEY; bool table_any = false;            // the bunch the warp's table holds; whether any of its queries is seeded
	const uint64_t nchunk = (A.nwork + A.chunk - 1) / A.chunk;
	LaneSeeds LS; LS.n = 0;                                                // local memory; touched only by lanes that see several queries or far-apart seeds

	auto key_of = [&](uint64_t i) -> unsigned long long {                  // (query0, nq) of work item i
		uint64_t r; uint32_t c, q0, n; get_work(A.W, i, r, c, q0, n);
		return ((unsigned long long)q0 << 8) | n;
	};

	for (uint64_t ch = (uint64_t)blockIdx.x * SEEDW_WARPS + warp; ch < nchunk; ch += (uint64_t)gridDim.x * SEEDW_WARPS) {
		uint64_t i = ch * A.chunk;
		const uint64_t cend = min(A.nwork, i + A.chunk);
		// ---- runs at the head of the chunk that continue a bunch begun in an earlier chunk belong to that chunk's warp ----
		if (i % SEEDW_SPLIT) {
			unsigned long long kprev = key_of(i - 1);
			for (;;) {
				const uint64_t me = i + lane;
				const unsigned long long k = me < cend ? key_of(me) : NOKEY;
				unsigned long long pk = __shfl_up_sync(FULL, k, 1);
				if (lane == 0) pk = kprev;
				const uint32_t b = __ballot_sync(FULL, k != pk || me % SEEDW_SPLIT == 0 || me >= cend);
				if (b) { i += __ffs(b) - 1; break; }
				kprev = __shfl_sync(FULL, k, 31); i += 32;
			}
			if (i >= cend) continue;
		}
		// ---- segments: <= 32 consecutive runs of one bunch ----
		for (;;) {
			uint64_t d_r = 0; uint32_t d_c = 0, q0 = 0, n = 0; bool d_ok = false;
			unsigned long long k = NOKEY;
			if (i + lane < A.nwork) { d_ok = get_work(A.W, i + lane, d_r, d_c, q0, n); k = ((unsigned long long)q0 << 8) | n; }
			const unsigned long long K = __shfl_sync(FULL, k, 0);
			const uint32_t lim = (uint32_t)min((uint64_t)32, min(A.nwork - i, (uint64_t)SEEDW_SPLIT - i % SEEDW_SPLIT));
			const uint32_t same = __ballot_sync(FULL, k == K && lane < lim);
			const uint32_t seglen = same == FULL ? 32u : (uint32_t)__ffs(~same) - 1u;        // >= 1
			d_ok = d_ok && lane < seglen;
			uint4 d_m = make_uint4(0, 0, 0, 0);
			if (d_ok) d_m = __ldg((const uint4 *)(A.meta + d_c));              // clump record of this lane's run: offset (uint4 units), length, flags
			const uint32_t d_r32 = (uint32_t)d_r;

			// ---- the bunch's window set ----
			if (K != tableK) {
				const uint32_t bq0 = (uint32_t)(K >> 8), bn = (uint32_t)(K & 255);
				for (uint32_t w = lane * 4; w < BW + A.hslots; w += 128) *(uint4 *)(bits + w) = make_uint4(0, 0, 0, 0);   // bitmap and slots are adjacent
				bool act = false; QInfo Q; Q.len = 0; Q.k = 0; Q.off = 0;
				if (l < bn) { Q = A.qi[bq0 + l]; act = (Q.cls & 1) != 0; }           // both half-warps read the same 16 records
				if (half == 0) kq[l] = Q.k;
				__syncwarp();
				const uint32_t np = act ? Q.k + 1u : 0u, plen = act ? Q.len / np : 0u;
				const uint32_t *Wq = A.qnib + (Q.off >> 3) + 3ull * (bq0 + l);
				for (uint32_t p = half; p < NPM; p += 2) {                     // thread = (query l, stretches of its half's parity)
					const uint32_t si = l * NPM + p;
					if (p < np) {
						const QStretch S = stretch_of(Wq, plen, p);
						*(uint4 *)(str + si * 4) = make_uint4(S.r0, S.r1, S.r2, S.E);
						#pragma unroll
						for (uint32_t j = 0; j < (uint32_t)STRIDE; ++j) {
							const QWin w = window_of(S, j);
							const uint32_t hv = seed_hash(w.kn, w.ko & HM);
							atomicOr(&bits[hv >> SHB], (0x80000000u >> (hv & 31)) | (FB >= 2 ? 0x80000000u >> ((hv >> 5) & 31) : 0u) | (FB >= 3 ? 0x80000000u >> ((hv >> 10) & 31) : 0u));
							const uint32_t e = si * STRIDE + j;
							nxt[e] = (uint16_t)atomicExch(&slots[(hv >> 10) & HSM], e + 1u);
						}
					} else *(uint4 *)(str + si * 4) = make_uint4(0, 0, 0, 0);
				}
				table_any = __any_sync(FULL, act);
				tableK = K;
				__syncwarp();
			}

			if (table_any && __any_sync(FULL, d_ok)) {
				// ---- scan: each half-warp walks every other run of the segment; item = NCH chunks (32 columns each) of a run ----
				struct RunD { uint64_t off; uint32_t len, flags, r; bool valid; };
				auto fetch = [&](uint32_t t) -> RunD {                        // record of run t of the segment, from the lane that holds it
					const uint32_t src = min(t, 31u);
					RunD D;
					D.off = (uint64_t)__shfl_sync(FULL, d_m.x, src) | ((uint64_t)__shfl_sync(FULL, d_m.y, src) << 32);
					D.len = __shfl_sync(FULL, d_m.z, src); D.flags = __shfl_sync(FULL, d_m.w, src); D.r = __shfl_sync(FULL, d_r32, src);
					D.valid = __shfl_sync(FULL, (uint32_t)d_ok, src) != 0 && t < seglen;
					return D;
				};
				// bulk copy of item g of run D into staging buffer b of this half-warp (its first lane issues; the others only wait later)
				auto stage = [&](const RunD &D, uint32_t g, uint32_t b) {
					if (D.valid && l == 0) {
						const uint32_t nchunks = (D.len + 31) >> 5, bytes = min((uint32_t)NCH, nchunks - g * NCH) * 256u;
						mbar_expect_tx(bar_s + 8 * b, bytes);
						bulk_g2s(stg_s + b * ITEM, A.db + D.off + (size_t)g * NCH * 16, bytes, bar_s + 8 * b);
					}
				};
				uint32_t ct = half, cg = 0;                                    // current run (index in the segment) and item within it
				RunD C = fetch(ct);
				uint32_t prev = 0, prev2 = 0, ambp = 0, ambp2 = 0;             // scan state carried from item to item of a run
				uint32_t iprev = 0, iprev2 = 0;                                // the two words before the current item (for verifying its first words)
				uint32_t sn = 0; bool serial = false;                          // seeds of the current run: sn 0 = in the lane's shared-memory record (one query, a hull), 2 = list (LS); serial: the lane verifies its own words
				bool more = true;

				auto seed = [&](uint32_t q, int dg) {                          // list form (lanes that see several queries, IUPAC clumps)
					if (sn == 0) {                                             // what the helpers gathered so far is one cluster: its two ends stand for it in the list
						const uint32_t q0 = ost[lane * 4];
						LS.n = 0; sn = 2;
						if (q0 != NOQ) { LS.q[0] = LS.q[1] = q0; LS.d[0] = (int)ost[lane * 4 + 1]; LS.d[1] = (int)ost[lane * 4 + 2]; LS.n = 2; }
					}
					if (LS.n < SEED_LIST) { LS.q[LS.n] = q; LS.d[LS.n] = dg; ++LS.n; }
					else lane_seeds_overflow(LS, q, dg);
				};

				auto step = [&]() {
					// ---- the item after this one: load it now, probe it in the next step ----
					const uint32_t ngroups = C.valid ? (((C.len + 31) >> 5) + NCH - 1) / NCH : 1u;
					const bool lastg = cg + 1 >= ngroups;
					const uint32_t nt = lastg ? ct + 2 : ct, ng = lastg ? 0u : cg + 1;
					const RunD F = fetch(nt);
					const RunD N = lastg ? F : C;
					__syncwarp();                                              // everyone is done with the other buffer (read in the previous step)
					stage(N, ng, cb ^ 1);
					bool emit = false; Surv ev; ev.task = 0; ev.lo = 0; ev.w_lane = 0; ev.scratch = 0;
					uint32_t m8 = 0, m4 = 0;                                   // flagged words of this lane's item: word w -> bit NW-1-w
					const bool amb_on = C.valid && (C.flags & ambsel) != 0;
					if (C.valid) {
						const uint32_t nchunks = (C.len + 31) >> 5;
						if (cg == 0) { prev = prev2 = ambp = ambp2 = 0; sn = 0; serial = amb_on; *(uint4 *)(ost + lane * 4) = make_uint4(NOQ, 0x7FFFFFFFu, 0x80000000u, 0u); }
						iprev = prev; iprev2 = prev2;
						const uint32_t sb = stg_s + cb * ITEM + l * 16;              // this lane's pieces: chunk c at sb + c * 256
						mbar_wait(bar_s + 8 * cb, (phase >> cb) & 1u); phase ^= 1u << cb;
						auto scan = [&](auto amb_c) {
							constexpr bool AMB = decltype(amb_c)::value;
							#pragma unroll
							for (int c = 0; c < NCH; ++c) {
								const uint4 cw = lds128(sb + c * 256);
								const uint32_t ws[4] = {cw.x, cw.y, cw.z, cw.w};
								#pragma unroll
								for (int j = 0; j < 4; ++j) {
									const uint32_t cu = ws[j];
									const uint32_t h8 = seed_hash(cu, prev & HM);
									const uint32_t f8 = lds32(bits_s + ((h8 >> SHB) << 2));
									uint32_t t8 = __funnelshift_l(0u, f8, h8);                                    // the window's bit(s) -> bit 31
									if (FB >= 2) t8 &= __funnelshift_l(0u, f8, h8 >> 5);
									if (FB >= 3) t8 &= __funnelshift_l(0u, f8, h8 >> 10);
									uint32_t t4 = 0;
									if (STRIDE == 4) {
										const uint32_t h4 = seed_hash(__funnelshift_r(prev, cu, 16), __funnelshift_r(prev2, prev, 16) & HM);
										const uint32_t f4 = lds32(bits_s + ((h4 >> SHB) << 2));
										t4 = __funnelshift_l(0u, f4, h4);
										if (FB >= 2) t4 &= __funnelshift_l(0u, f4, h4 >> 5);
										if (FB >= 3) t4 &= __funnelshift_l(0u, f4, h4 >> 10);
									}
									if (AMB) {
										const uint32_t ambc = amb_nibbles(cu, ADD);
										if (ambc | ambp) t8 = 0x80000000u;
										if (ambc | ambp | ambp2) t4 = 0x80000000u;
										ambp2 = ambp; ambp = ambc;
									}
									m8 = __funnelshift_l(t8, m8, 1);                       // word w of the item -> bit NW-1-w
									if (STRIDE == 4) m4 = __funnelshift_l(t4, m4, 1);
									prev2 = prev; prev = cu;
								}
							}
						};
						if (amb_on) scan(std::true_type{}); else scan(std::false_type{});
						const uint32_t nvalid = min((uint32_t)NW, nchunks * 4 - cg * NW);
						const uint32_t vm = nvalid >= 32 ? FULL : (((1u << nvalid) - 1u) << (NW - nvalid));
						m8 &= vm; m4 &= vm;
					}
					// ---- verification of the flagged words against the window table; seeds -> clusters ----
					// A flagged word is rare per lane (a false positive of the filter, or the one or two windows of a true alignment), so a lane
					// that verified its own would keep the other 31 waiting.  Instead the flagged words of the whole warp go into a queue in
					// shared memory and are verified 32 at a time, one per HELPER lane (the owner's words are in its staging buffer, readable by
					// anyone); a match returns to its owner by shuffle.
					if (__any_sync(FULL, (m8 | m4) != 0u)) {
						const uint32_t sb = stg_s + cb * ITEM + l * 16;
						// the owner's own (serial) verification: IUPAC clumps, lanes with seeds of several queries, a full queue
						auto own = [&](uint32_t o8, uint32_t o4) {
							if (o8 | o4) serial = true;                               // its seeds now live in the list: later items of the run must go there too
							auto word_at = [&](int wl) -> uint32_t { return wl >= 0 ? lds32(sb + (uint32_t)(wl >> 2) * 256 + (uint32_t)(wl & 3) * 4) : (wl == -1 ? iprev : iprev2); };
							auto verify = [&](uint32_t wi, int e) {
								const int wl = (int)(wi - cg * NW);
								const uint32_t cu = word_at(wl), pv = word_at(wl - 1), pv2 = e == 4 ? word_at(wl - 2) : 0u;
								const uint32_t rn = e == 8 ? cu : __funnelshift_r(pv, cu, 16), ro = (e == 8 ? pv : __funnelshift_r(pv2, pv, 16)) & HM;
								const int x1 = (int)(wi * 8 + e);
								if (amb_on && (amb_nibbles(rn, ADD) | amb_nibbles(ro, ADD))) {       // IUPAC codes in the window: every window of the bunch, through the table
									for (uint32_t si = 0; si < 16 * NPM; ++si) {
										const uint4 rec = *(const uint4 *)(str + si * 4);
										if (!rec.w) continue;
										QStretch S; S.r0 = rec.x; S.r1 = rec.y; S.r2 = rec.z; S.E = rec.w;
										for (uint32_t j = 0; j < (uint32_t)STRIDE; ++j) {
											const QWin w = window_of(S, j);
											if (window_matches_table(sM, w.kn, w.ko & HM, rn, ro, A.SL.w)) seed(((si * NPI) >> 16), x1 - (int)w.y1);
										}
									}
								} else {
									const uint32_t hv = seed_hash(rn, ro);
									for (uint32_t en = slots[(hv >> 10) & HSM]; en; en = nxt[en - 1]) {
										const uint32_t si = (en - 1) / STRIDE, j = (en - 1) % STRIDE;
										const uint4 rec = *(const uint4 *)(str + si * 4);
										QStretch S; S.r0 = rec.x; S.r1 = rec.y; S.r2 = rec.z; S.E = rec.w;
										const QWin w = window_of(S, j);
										if (w.kn == rn && (w.ko & HM) == ro) seed(((si * NPI) >> 16), x1 - (int)w.y1);
									}
								}
							};
							while (o8 | o4) {
								const uint32_t b = 31 - __clz(o8 | o4), bit = 1u << b, wi = cg * NW + (NW - 1 - b);
								if (STRIDE == 4 && (o4 & bit)) verify(wi, 4);
								if (o8 & bit) verify(wi, 8);
								o8 &= ~bit; o4 &= ~bit;
							}
						};
						const uint32_t s8 = m8, s4 = m4;                           // kept: a lane whose helpers report a second query redoes the item itself
						if (serial) { own(m8, m4); m8 = m4 = 0; }
						if (VM == 1) {
							// ---- one pass per flagged lane: the 32 lanes take the (<= 32) words of the owner's item, each verifies its word if flagged; the matches
							//      of the pass (query range, diagonal range) come back by warp reduction and the owner folds them into its record -- no queue, no
							//      shared-memory atomics.  A record stays ONE cluster of ONE query exactly as below: on a second query or seeds out of reach the
							//      record is left as it was and the owner redoes the item itself, into a proper list.
							uint32_t fl = __ballot_sync(FULL, (m8 | m4) != 0u);
							while (fl) {
								const uint32_t owner = (uint32_t)__ffs(fl) - 1u; fl &= fl - 1u;
								const uint32_t o8 = __shfl_sync(FULL, m8, owner), o4 = STRIDE == 4 ? __shfl_sync(FULL, m4, owner) : 0u;
								const uint32_t o_cg = __shfl_sync(FULL, cg, owner), o_ip = __shfl_sync(FULL, iprev, owner), o_ip2 = __shfl_sync(FULL, iprev2, owner);
								const uint32_t bit = lane < (uint32_t)NW ? 1u << (NW - 1 - lane) : 0u;
								uint32_t qlo = NOQ, qhi = 0u; int dlo = 0x7FFFFFFF, dhi = (int)0x80000000;
								if ((o8 | o4) & bit) {
									const uint32_t osb = stgw_s + (owner >> 4) * 2 * ITEM + cb * ITEM + (owner & 15) * 16;
									auto word = [&](int w) -> uint32_t { return w >= 0 ? lds32(osb + (uint32_t)(w >> 2) * 256 + (uint32_t)(w & 3) * 4) : (w == -1 ? o_ip : o_ip2); };
									const uint32_t cu = word((int)lane), pv = word((int)lane - 1), pv2 = (STRIDE == 4 && (o4 & bit)) ? word((int)lane - 2) : 0u;
									#pragma unroll
									for (int e = (STRIDE == 4 ? 4 : 8); e <= 8; e += 4) {
										if (!((e == 4 ? o4 : o8) & bit)) continue;
										const uint32_t rn = e == 4 ? __funnelshift_r(pv, cu, 16) : cu, ro = (e == 4 ? __funnelshift_r(pv2, pv, 16) : pv) & HM;
										const int x1 = (int)((o_cg * NW + lane) * 8 + (uint32_t)e);
										for (uint32_t en = slots[(seed_hash(rn, ro) >> 10) & HSM]; en; en = nxt[en - 1]) {
											const uint32_t si = (en - 1) / STRIDE, j = (en - 1) % STRIDE;
											const uint4 rec = *(const uint4 *)(str + si * 4);
											QStretch S; S.r0 = rec.x; S.r1 = rec.y; S.r2 = rec.z; S.E = rec.w;
											const QWin w = window_of(S, j);
											if (w.kn == rn && (w.ko & HM) == ro) {
												const uint32_t q = (si * NPI) >> 16; const int dg = x1 - (int)w.y1;
												qlo = min(qlo, q); qhi = max(qhi, q); dlo = min(dlo, dg); dhi = max(dhi, dg);
											}
										}
									}
								}
								const uint32_t Qlo = __reduce_min_sync(FULL, qlo);
								if (Qlo != NOQ) {
									const uint32_t Qhi = __reduce_max_sync(FULL, qhi);
									const int Dlo = __reduce_min_sync(FULL, dlo), Dhi = __reduce_max_sync(FULL, dhi);
									if (lane == owner) {
										const uint4 rec = *(const uint4 *)(ost + lane * 4);
										const int nlo = min((int)rec.y, Dlo), nhi = max((int)rec.z, Dhi);
										if (Qlo != Qhi || (rec.x != NOQ && rec.x != Qlo) || nhi - nlo > 2 * (int)kq[Qlo] + 1) own(s8, s4);
										else *(uint4 *)(ost + lane * 4) = make_uint4(Qlo, (uint32_t)nlo, (uint32_t)nhi, 0u);
									}
								}
								__syncwarp();
							}
							m8 = m4 = 0;
						}
						uint32_t cnt = (uint32_t)(__popc(m8) + __popc(m4));
						if (VM == 0 && __reduce_add_sync(FULL, cnt) > HQ) { own(m8, m4); m8 = m4 = 0; cnt = 0; }       // more flagged words than the queue holds (rare): everyone its own
						if (VM == 0 && __any_sync(FULL, cnt != 0u)) {
							// ---- enqueue ----
							if (lane == 0) *hqn = 0;
							__syncwarp();
							const uint4 before = *(const uint4 *)(ost + lane * 4);
							uint32_t pos = cnt ? atomicAdd(hqn, cnt) : 0u;
							while (m8 | m4) {
								const uint32_t b = 31 - __clz(m8 | m4), bit = 1u << b;
								const bool four = STRIDE == 4 && (m4 & bit);           // the half-word window of a word comes before its full-word window
								hq[pos++] = (lane << 8) | (four ? 128u : 0u) | (uint32_t)(NW - 1 - b);
								if (four) m4 &= ~bit; else m8 &= ~bit;
							}
							__syncwarp();
							const uint32_t total = *hqn;
							// ---- helpers: one queue entry per lane and round; a match goes into the owner's record ----
							for (uint32_t base = 0; base < total; base += 32) {
								const bool have = base + lane < total;
								const uint32_t ent = have ? hq[base + lane] : 0u, owner = ent >> 8, wl = ent & 127u;
								const bool four = (ent & 128u) != 0;
								const uint32_t o_cg = __shfl_sync(FULL, cg, owner), o_ip = __shfl_sync(FULL, iprev, owner), o_ip2 = __shfl_sync(FULL, iprev2, owner);
								if (have) {
									const uint32_t osb = stgw_s + (owner >> 4) * 2 * ITEM + cb * ITEM + (owner & 15) * 16;
									auto word = [&](int w) -> uint32_t { return w >= 0 ? lds32(osb + (uint32_t)(w >> 2) * 256 + (uint32_t)(w & 3) * 4) : (w == -1 ? o_ip : o_ip2); };
									const uint32_t cu = word((int)wl), pv = word((int)wl - 1), pv2 = four ? word((int)wl - 2) : 0u;
									const uint32_t rn = four ? __funnelshift_r(pv, cu, 16) : cu, ro = (four ? __funnelshift_r(pv2, pv, 16) : pv) & HM;
									const int x1 = (int)((o_cg * NW + wl) * 8 + (four ? 4 : 8));
									for (uint32_t en = slots[(seed_hash(rn, ro) >> 10) & HSM]; en; en = nxt[en - 1]) {
										const uint32_t si = (en - 1) / STRIDE, j = (en - 1) % STRIDE;
										const uint4 rec = *(const uint4 *)(str + si * 4);
										QStretch S; S.r0 = rec.x; S.r1 = rec.y; S.r2 = rec.z; S.E = rec.w;
										const QWin w = window_of(S, j);
										if (w.kn == rn && (w.ko & HM) == ro) {
											const uint32_t q = ((si * NPI) >> 16), was = atomicCAS(&ost[owner * 4], NOQ, q);
											const int dg = x1 - (int)w.y1;
											if (was == NOQ || was == q) { atomicMin((int *)&ost[owner * 4 + 1], dg); atomicMax((int *)&ost[owner * 4 + 2], dg); }
											else ost[owner * 4 + 3] = 1u;                    // a second query on this lane: the owner takes over
										}
									}
								}
							}
							__syncwarp();
							// A record stays ONE cluster of ONE query: if the helpers met a second query, or seeds of this item lie out of reach
							// (2k+1 diagonals) of the cluster, the owner puts the record back as it was before the item and redoes the item itself,
							// into a proper list (far seeds = separate clusters, each its own narrow band in k_extend)
							if (cnt) {
								const uint4 now = *(const uint4 *)(ost + lane * 4);
								if (now.w || (now.x != NOQ && (int)now.z - (int)now.y > 2 * (int)kq[now.x] + 1)) {
									*(uint4 *)(ost + lane * 4) = before;
									own(s8, s4);
								}
							}
						}
					}
					if (C.valid) {
						// ---- end of the run: its seeds leave as survivors ----
						if (lastg) {
							const uint32_t task0 = (C.r + A.W.run_base) * BG_RUN_MAX;
							if (sn == 2) { lane_seeds_emit(LS, kq, task0, l, A.surv, A.surv_cap, A.counters); sn = 0; }
							else {
								const uint4 st = *(const uint4 *)(ost + lane * 4);
								if (st.x != NOQ) {
									const int k0 = (int)kq[st.x], dlo = (int)st.y, dhi = (int)st.z;
									const uint32_t W = (uint32_t)(dhi - dlo + 2 * k0 + 1);
									emit = true; ev.task = task0 + st.x; ev.lo = dlo - k0; ev.w_lane = (W << 8) | (1u << 4) | l;
									if (W > 64) atomicMax(&A.counters[C_SCRATCH], W);
								}
							}
						}
					}
					const uint32_t em = __ballot_sync(FULL, emit);
					if (em) {
						if (scount + __popc(em) > 32) flush();
						if (emit) sbuf[scount + __popc(em & ((1u << lane) - 1u))] = make_uint4(ev.task, (uint32_t)ev.lo, ev.w_lane, ev.scratch);
						scount += __popc(em);
						__syncwarp();
					}
					ct = nt; cg = ng; C = N; cb ^= 1;
					more = __any_sync(FULL, ct < seglen);
				};

				__syncwarp();
				stage(C, 0, cb);
				while (more) step();
			}

			i += seglen;
			if (i >= A.nwork) break;
			if (i >= cend && (i % SEEDW_SPLIT == 0 || key_of(i) != K)) break;  // past the chunk: go on only while the bunch does
		}
	}
	if (scount) flush();
}

// ---------------------------------------------------------------------------------------------
// Phase A2: Myers bit-parallel prefix filter.  128 threads = 8 tasks x 16 lanes.
// ---------------------------------------------------------------------------------------------
struct FilterArgs {
	const uint4 *db; const uint64_t *clump_off; const uint32_t *clump_len;
	const QInfo *qi; const uint8_t *codes; const uint32_t *Sterm; Work W;
	Surv *surv; uint32_t surv_cap; uint32_t *counters;
	uint32_t c16;        // the constant 16, passed at run time so ptxas keeps IMAD.HI (FMA pipe)
	const uint32_t *todo; // number of queries k_seed did not take (device; NULL = unknown, run)
};

// Hyyro's block formulation of Myers' bit-vector algorithm over the first P rows of the query, NW words of 32 rows: the carry of
// the addition and the shifted horizontal deltas pass from one word to the next as (hp, hm); the last word's pair is the change of
// the row-P value from one column to the next.  P is chosen per query (filt_code): about four times its budget, so that a
// prefix within budget is rare by chance -- with 32 rows a budget of 15 is met almost everywhere and the filter passes everything.
// Seed columns (row-P value <= k) give diagonals x - P; consecutive ones merge into clusters [d - k, d + k] (at most CLUS_MAX per
// lane, the closest fused beyond that), each cluster one survivor -- not one hull over the whole clump, which on the long
// references of a sheared genome database would be a band as wide as the clump.
__device__ __forceinline__ uint32_t mulhi(uint32_t a, uint32_t b) {      // nibble extraction on the FMA pipe (the ALU pipe is the kernel's bound)
	uint32_t d; asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
template <int NW>
__global__ void __launch_bounds__(128) k_filter(FilterArgs A) {
	__shared__ uint32_t sPeq[8][NW][16];
	if (A.todo && *A.todo == 0) return;                           // every query of the batch went to k_seed
	const uint32_t slot = threadIdx.x >> 4, lane = threadIdx.x & 15;
	const uint64_t ngroups = A.W.nruns * 2;                       // 8 task ids per group
	for (uint64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
	__syncwarp();                                                 // the previous group's readers are done with sPeq
	const uint64_t t = grp * 8 + slot;                            // task id = run * 16 + query-in-run
	const uint64_t r = t >> 4; const uint32_t qi_ = (uint32_t)t & 15;
	bool valid = r < A.W.nruns;
	uint32_t q = 0, c = 0, q0 = 0, n = 0;
	if (valid) { valid = get_run(A.W, r, c, q0, n) && qi_ < n; q = q0 + qi_; }
	QInfo Q; Q.cls = 1; Q.P = 0; Q.k = 0; Q.off = 0;
	if (valid) { Q = A.qi[q]; valid = !(Q.cls & 1) && filt_words(Q.P) == (uint32_t)NW; }   // seed-eligible queries were handled by k_seed; other prefix lengths by the other instances
	const int P = (int)filt_rows(Q.P), k = Q.k;
	const int sh = NW == 1 ? 32 - P : 0;                           // a prefix shorter than a word sits in its top bits
	{	// the match masks of reference code `lane` against the prefix rows (0 = the scoring table says the pair costs nothing)
		const uint8_t *qs = A.codes + Q.off;
		#pragma unroll 1
		for (int w = 0; w < NW; ++w) {
			uint32_t m = 0;
			if (valid) {
				if (NW == 1 && P < 32) m = (1u << sh) - 1u;
				const int rows = min(32, P - 32 * w);
				for (int y = 0; y < rows; ++y) if (A.Sterm[(qs[32 * w + y] & 15) * 16 + lane] == 0) m |= 1u << (y + sh);
			}
			sPeq[slot][w][lane] = m;
		}
	}
	__syncwarp();
	if (!valid) continue;
	const uint32_t L = A.clump_len[c];
	const uint4 *base = A.db + A.clump_off[c] + lane;
	const uint32_t eqs = (uint32_t)__cvta_generic_to_shared(&sPeq[slot][0][0]);

	uint32_t Pv[NW], Mv[NW];
	#pragma unroll
	for (int w = 0; w < NW; ++w) { Pv[w] = ~0u; Mv[w] = 0; }
	if (NW == 1 && P < 32) Pv[0] = ~0u << (32 - P);
	int score = P;                                                 // row-P value of the column before the first
	int clo = 0, chi = INT32_MIN; bool open = false; Clus C; C.n = 0;
	const uint32_t nwords = (L + 7) >> 3;
	uint4 w4 = base[0];
	for (uint32_t wi0 = 0; wi0 < nwords; wi0 += 4) {
		uint4 nx = w4;
		if (wi0 + 4 < nwords) nx = base[(size_t)((wi0 >> 2) + 1) * 16];
		const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
		#pragma unroll 1
		for (int wi = 0; wi < 4; ++wi) {
			if (wi0 + wi >= nwords) break;
			const uint32_t w = ws[wi];
			#pragma unroll
			for (int j = 0; j < 8; ++j) {
				const uint32_t code = mulhi(j == 7 ? w : w << (28 - 4 * j), A.c16);
				uint32_t hp = 0, hm = 0;                                 // row 0 is all zero: nothing enters the first word
				#pragma unroll
				for (int b = 0; b < NW; ++b) {
					uint32_t Eq;
					asm volatile("ld.shared.u32 %0, [%1];" : "=r"(Eq) : "r"(eqs + (uint32_t)b * 64u + code * 4u));
					const uint32_t Xv = Eq | Mv[b];
					Eq |= hm;
					const uint32_t Xh = (((Eq & Pv[b]) + Pv[b]) ^ Pv[b]) | Eq;
					uint32_t Ph = Mv[b] | ~(Xh | Pv[b]), Mh = Pv[b] & Xh;
					const uint32_t hpo = Ph >> 31, hmo = Mh >> 31;
					Ph = (Ph << 1) | hp; Mh = (Mh << 1) | hm;
					Pv[b] = Mh | ~(Xv | Ph);
					Mv[b] = Ph & Xv;
					hp = hpo; hm = hmo;
				}
				score += (int)hp - (int)hm;
				if (score <= k) {                                        // seed: D[P][x] <= k
					const int x = (int)((wi0 + wi) * 8 + j) + 1;
					if (x <= (int)L) {
						const int d = x - P;
						if (open && d - k <= chi + 1) chi = d + k;
						else { if (open) clus_add(C, clo, chi); clo = d - k; chi = d + k; open = true; }
					}
				}
			}
		}
		w4 = nx;
	}
	if (open) clus_add(C, clo, chi);
	emit_clusters(C, (uint32_t)t + A.W.run_base * BG_RUN_MAX, lane, A.surv, A.surv_cap, A.counters);
	}
}

// ---------------------------------------------------------------------------------------------
// Phase B: exact banded DP with the pass-2 triple.
// ---------------------------------------------------------------------------------------------
#define NCLASS 9
struct ExtendArgs {
	const uint32_t *dbw; const ClumpMeta *meta;
	const uint8_t *codes; const uint32_t *qnib; const QInfo *qi; Work W;
	const Surv *surv; uint32_t surv_cap; const uint32_t *counters;
	const uint32_t *cls; const uint4 *xs;                // survivors of this launch binned by band class, as expanded records (k_bin_*)
	uint32_t np_stage[NCLASS], qp_stage;                 // staging slot per thread and class: reference pieces, query word quads (uint4 each; 0 = read global memory)
	Res *res; uint32_t *best; const uint32_t *Sterm;   // Sterm[q*16+r] = S << 22
	uint32_t gen_mode;                                    // experiments on the generic band path (BURST_B200_GEN_MODE)
	uint32_t smem_w;                                      // generic bands of up to smem_w cells live in shared memory (cell d of thread t at word d * 128 + t)
	uint32_t *scratch; uint32_t scratch_w;           // generic bands (wider than 64): scratch_w cells per thread of the generic launch, cell d of thread t at scratch[d * threads + t]
	unsigned long long *band_cells;
	int mode;
};

__device__ __forceinline__ uint32_t fetch_code(const uint32_t *lanew, uint32_t xi, uint32_t L) {
	if (xi >= L) return 0;                                   // also catches "negative" columns
	const uint32_t w = __ldg(lanew + (size_t)(xi >> 5) * 64 + ((xi >> 3) & 3));
	return (w >> ((xi & 7) * 4)) & 15;
}

// One DP cell.  diag/up are row y-1, left is row y; sterm = S(q[y], r[x]) << 22.
__device__ __forceinline__ uint32_t cell(uint32_t diag, uint32_t up, uint32_t left, uint32_t sterm, uint32_t inf) {
	uint32_t t = viaddmin(up, KEY_UP, diag + sterm);         // burst.c:767-780
	t = viaddmin(left, KEY_LEFT, t);                         // burst.c:783-799
	return min(t, inf) & KEY_CLEAR;                          // burst.c:802-803
}

// Eight consecutive codes of one lane starting at (0-based, possibly negative) column cb, one nibble each;
// columns outside the clump read as 0 (pad).  Two 32-bit loads; consecutive groups share one.
__device__ __forceinline__ uint32_t lane_word_or0(const uint32_t *lanew, int wi, int nwords) {
	return (wi >= 0 && wi < nwords) ? __ldg(lanew + (size_t)(wi >> 2) * 64 + (wi & 3)) : 0u;
}

// Band classes: a survivor goes to the narrowest register band that holds its cluster (W = hull of the seed
// diagonals + 2k); above 64 the band lives in global scratch.  Survivors are binned by class before the sweep
// (k_bin_*), so that a warp's 32 threads run the same instantiation on 32 survivors.
__host__ __device__ __forceinline__ constexpr int class_width(int c) { return c == 0 ? 5 : c == 1 ? 8 : c == 2 ? 12 : c == 3 ? 16 : c == 4 ? 24 : c == 5 ? 32 : c == 6 ? 48 : c == 7 ? 64 : 0; }
__device__ __forceinline__ uint32_t class_of(uint32_t W) { return W <= 5 ? 0u : W <= 8 ? 1u : W <= 12 ? 2u : W <= 16 ? 3u : W <= 24 ? 4u : W <= 32 ? 5u : W <= 48 ? 6u : W <= 64 ? 7u : 8u; }
// cls: [0, NCLASS) counts, [16, 16+NCLASS) first position in `order`, [32, 32+NCLASS) fill cursors
__global__ void k_bin_count(const Surv *__restrict__ surv, const uint32_t *__restrict__ counters, uint32_t surv_cap, const uint32_t *__restrict__ firstp, uint32_t *__restrict__ cls) {
	__shared__ uint32_t sh[NCLASS];
	if (threadIdx.x < NCLASS) sh[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t nsurv = min(counters[C_SURV], surv_cap), first = firstp ? min(*firstp, nsurv) : 0u;
	for (uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x; i < nsurv; i += gridDim.x * blockDim.x) atomicAdd(&sh[class_of(surv[i].w_lane >> 8)], 1u);
	__syncthreads();
	if (threadIdx.x < NCLASS && sh[threadIdx.x]) atomicAdd(&cls[threadIdx.x], sh[threadIdx.x]);
}
__global__ void k_bin_offsets(const uint32_t *__restrict__ counters, uint32_t surv_cap, const uint32_t *__restrict__ firstp, uint32_t *__restrict__ cls) {
	if (threadIdx.x) return;
	const uint32_t nsurv = min(counters[C_SURV], surv_cap);
	uint32_t o = firstp ? min(*firstp, nsurv) : 0u;
	for (int c = 0; c < NCLASS; ++c) { cls[16 + c] = o; cls[32 + c] = 0; o += cls[c]; }
}
// One record per survivor, in class order, holding everything the sweep needs: the survivor itself plus what the chain
// survivor -> run -> (query record, clump record) resolves to.  This kernel has a thread per survivor and nothing to compute, so the
// chain's latency hides behind its parallelism; k_extend then starts from ONE coalesced 48-byte read.
struct XSurv {
	uint32_t task; int32_t lo; uint32_t w_lane, scratch;      // the survivor
	uint32_t coff_lo, coff_hi, L, qw;                           // clump: offset in uint4 units, length; query: word index of its packed bases in qnib
	uint32_t m, slot, k_flags, index;                           // query length, slot, budget | plain << 16; index of the survivor in the list
};
__global__ void k_bin_scatter(const Surv *__restrict__ surv, const uint32_t *__restrict__ counters, uint32_t surv_cap, const uint32_t *__restrict__ firstp,
		uint32_t *__restrict__ cls, uint4 *__restrict__ xs, Work W, const QInfo *__restrict__ qi, const ClumpMeta *__restrict__ meta) {
	const uint32_t nsurv = min(counters[C_SURV], surv_cap), first = firstp ? min(*firstp, nsurv) : 0u;
	const uint32_t lane = threadIdx.x & 31;
	for (uint32_t i0 = first + (blockIdx.x * blockDim.x + threadIdx.x - lane); i0 < nsurv; i0 += gridDim.x * blockDim.x) {
		const uint32_t i = i0 + lane;
		Surv sv; sv.task = 0; sv.lo = 0; sv.w_lane = 0; sv.scratch = 0;
		if (i < nsurv) sv = surv[i];
		const uint32_t c = i < nsurv ? class_of(sv.w_lane >> 8) : 0xFFu;
		uint32_t pos = 0;
		// one atomic per class present in the warp; positions inside a class keep the survivor order
		for (uint32_t todo = __ballot_sync(0xFFFFFFFFu, i < nsurv); todo;) {
			const uint32_t lead = __ffs(todo) - 1, cc = __shfl_sync(0xFFFFFFFFu, c, lead);
			const uint32_t same = __ballot_sync(0xFFFFFFFFu, c == cc);
			uint32_t base = 0;
			if (lane == lead) base = atomicAdd(&cls[32 + cc], (uint32_t)__popc(same));
			base = __shfl_sync(0xFFFFFFFFu, base, lead);
			if (c == cc) pos = cls[16 + cc] + base + __popc(same & ((1u << lane) - 1u));
			todo &= ~same;
		}
		if (i < nsurv) {
			uint32_t cl, q0, n;
			get_run(W, (sv.task >> 4) - W.run_base, cl, q0, n);
			const uint32_t qix = q0 + (sv.task & 15);
			const uint4 cm = __ldg((const uint4 *)(meta + cl));
			const QInfo Q = qi[qix];
			xs[(size_t)pos * 3] = make_uint4(sv.task, (uint32_t)sv.lo, sv.w_lane, sv.scratch);
			xs[(size_t)pos * 3 + 1] = make_uint4(cm.x, cm.y, cm.z, (uint32_t)((Q.off >> 3) + 3ull * qix + 2));
			xs[(size_t)pos * 3 + 2] = make_uint4(Q.len, Q.slot, (uint32_t)Q.k | ((Q.cls & 2u) << 15), i);
		}
	}
}

// cp.async (LDGSTS): 16 bytes global -> shared without a register round trip; groups complete in order
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// The sweep of one survivor is ~100 warp instructions, a DRAM round trip is several thousand cycles: a thread that fetched its words when
// it needed them would idle almost all the time.  So every thread runs a three-deep pipeline over ITS survivors (p, p+S, p+2S, ..):
// the 48-byte record of survivor p+2S is loading into registers while the reference pieces and packed query words of survivor p+S
// travel into the thread's second staging slot in shared memory (cp.async), while survivor p is swept out of the first slot.
template <int WMAX>
__global__ void __launch_bounds__(128) k_extend(ExtendArgs A) {
	__shared__ uint32_t sS[256];
	extern __shared__ __align__(16) uint4 xstage[];
	constexpr int CLS = WMAX == 5 ? 0 : WMAX == 8 ? 1 : WMAX == 12 ? 2 : WMAX == 16 ? 3 : WMAX == 24 ? 4 : WMAX == 32 ? 5 : WMAX == 48 ? 6 : WMAX == 64 ? 7 : 8;
	const uint32_t begin = A.cls[16 + CLS], count = A.cls[CLS];
	if (!count) return;
	for (int i = threadIdx.x; i < 256; i += blockDim.x) sS[i] = A.Sterm[i];
	__syncthreads();
	unsigned long long cells = 0;
	constexpr int WBS = WMAX ? WMAX : 1;
	const uint32_t NPS = WMAX ? A.np_stage[CLS] : 0u, QPS = NPS ? A.qp_stage : 0u, SLOT = NPS + QPS;       // uint4 per staging slot: reference pieces, query words
	const uint32_t slot_s = (uint32_t)__cvta_generic_to_shared(xstage) + threadIdx.x * 2 * SLOT * 16;
	const uint32_t S = gridDim.x * blockDim.x;
	struct Staged { int c0w, c1w; uint32_t qsh; bool on; };                // staged reference words [c0w, c1w), query words start qsh words into the slot
	auto load_x = [&](uint32_t p, uint4 (&X)[3]) { X[0] = __ldg(A.xs + (size_t)(begin + p) * 3); X[1] = __ldg(A.xs + (size_t)(begin + p) * 3 + 1); X[2] = __ldg(A.xs + (size_t)(begin + p) * 3 + 2); };
	auto stage = [&](const uint4 (&X)[3], uint32_t b) -> Staged {
		Staged T; T.c0w = 0; T.c1w = 0; T.qsh = X[1].w & 3u; T.on = false;
		if (!SLOT) return T;
		const int lo = (int)X[0].y, L = (int)X[1].z, m = (int)X[2].x;
		const int fc = max(lo - 1, 0), lc = min(lo - 1 + m + WBS + 16, L - 1);         // columns (0-based) the sweep can touch
		const int c0 = fc >> 5, np = lc >= fc ? (lc >> 5) - c0 + 1 : 0;
		const uint32_t nq4 = (T.qsh + (uint32_t)((m + 7) >> 3) + 3u) >> 2;
		if ((uint32_t)np > NPS || nq4 > QPS) return T;                                  // does not fit the slot: this survivor reads global memory directly
		const uint32_t *src = A.dbw + ((uint64_t)X[1].x | ((uint64_t)X[1].y << 32)) * 4 + (X[0].z & 15u) * 4 + (size_t)c0 * 64;
		const uint32_t dst = slot_s + b * SLOT * 16;
		for (int j = 0; j < np; ++j) cp_async16(dst + j * 16, src + (size_t)j * 64);
		const uint32_t *qsrc = A.qnib + (X[1].w & ~3u);
		for (uint32_t j = 0; j < nq4; ++j) cp_async16(dst + (NPS + j) * 16, qsrc + j * 4);
		T.c0w = c0 * 4; T.c1w = (c0 + np) * 4; T.on = true;
		return T;
	};
	uint32_t p = blockIdx.x * blockDim.x + threadIdx.x, buf = 0;
	uint4 X0[3], X1[3], X2[3];
	bool v0 = p < count, v1 = v0 && p + S < count;
	Staged T0, T1; T0.on = T1.on = false; T0.c0w = T0.c1w = T1.c0w = T1.c1w = 0; T0.qsh = T1.qsh = 0;
	if (v0) { load_x(p, X0); T0 = stage(X0, 0); }
	cp_async_commit();
	if (v1) load_x(p + S, X1);
	for (; v0; p += S, buf ^= 1) {
		const uint32_t kbest = A.mode == BG_MODE_MIN ? __ldcg(A.best + X0[2].y) : 0xFFFFFFFFu;   // the slot's running minimum: asked for before the staging work below, used after it
		if (v1) T1 = stage(X1, buf ^ 1);
		cp_async_commit();
		const bool v2 = v1 && p + 2 * S < count;
		if (v2) load_x(p + 2 * S, X2);
		if (v1) cp_async_wait1(); else cp_async_wait0();                   // everything but the group just issued has landed: this survivor's slot is ready (an empty newest group completes at once and does not count)
		const uint4 x0 = X0[0], x1 = X0[1], x2 = X0[2];
		const Staged T = T0;
		const uint32_t ref_s = slot_s + buf * SLOT * 16, q_s = ref_s + NPS * 16 + T.qsh * 4;
		// rotate the pipeline registers now: the body below ends in `continue` on some paths
		X0[0] = X1[0]; X0[1] = X1[1]; X0[2] = X1[2]; X1[0] = X2[0]; X1[1] = X2[1]; X1[2] = X2[2]; T0 = T1; v0 = v1; v1 = v2;
		Surv sv; sv.task = x0.x; sv.lo = (int32_t)x0.y; sv.w_lane = x0.z; sv.scratch = x0.w;
		const uint32_t i = x2.w;
		const uint32_t W = sv.w_lane >> 8, lane = sv.w_lane & 15;
		const uint32_t m = x2.x, L = x1.z, slot = x2.y;
		const bool plain = (x2.z >> 16) & 1u;
		const uint32_t *lanew = A.dbw + ((uint64_t)x1.x | ((uint64_t)x1.y << 32)) * 4 + lane * 4;
		uint32_t k = x2.z & 0xFFFFu;
		k = min(k, kbest);
		uint32_t inf = (k + 1) << 22;
		const int lo = sv.lo;
		constexpr int WB = WMAX ? WMAX : 1;
		const int Wd = WMAX ? WMAX : (int)W;                 // cells per row actually swept
		uint32_t a[WB];                                      // band, register resident when WMAX > 0
		uint32_t *g = A.scratch + (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // generic path: this thread's band in global scratch, interleaved with the other threads'
		size_t GT = (size_t)gridDim.x * blockDim.x;
		if (WMAX == 0 && W <= A.smem_w) { g = (uint32_t *)xstage + threadIdx.x; GT = blockDim.x; }   // the usual case: a tenth of the latency per cell
		if (WMAX == 0 && W > A.scratch_w) { Res z; z.a = 0; z.b = 0; z.slot = slot; A.res[i] = z; continue; }   // (the host sees the width in the counter, grows the scratch and redoes the batch)
		bool dead = false, striped = false;
		uint32_t y = 1;
		uint32_t sbk = KEY_NONE >> 11, sbshr = 0, sfp = 0;       // last-row selection of the striped sweep

		if (WMAX) {
			// The band slides one column per row, so row y needs exactly one new reference code (column y+lo+WB-1)
			// and one query code: both are streamed as packed words, one 32-bit read of each per 8 rows
			// (the query from its nibble-packed copy, the lane from the DB pieces), out of the staging slot.
			constexpr int NW = (WB + 7) / 8;
			constexpr uint32_t TOPMASK = (WB & 7) ? (1u << (4 * (WB & 7))) - 1u : 0xFFFFFFFFu;   // nibbles of the last window word inside the band
			uint32_t win[NW];                                // codes of columns x0 .. x0+WB-1, one nibble each
			const int nwords = (int)((L + 7) >> 3);
			const uint32_t *Wq = A.qnib + x1.w;
			// word wi of the lane (0 outside the clump) / packed word g of the query
			auto refw = [&](int wi) -> uint32_t {
				if (T.on) return (wi >= T.c0w && wi < T.c1w) ? lds32(ref_s + (uint32_t)(wi - T.c0w) * 4) : 0u;
				return lane_word_or0(lanew, wi, nwords);
			};
			auto qwd = [&](uint32_t gg) -> uint32_t { return T.on ? lds32(q_s + gg * 4) : __ldg(Wq + gg); };
			// row 0: zero for columns 0..L (burst.c:4052 calloc / 723-725), absent elsewhere
			#pragma unroll
			for (int d = 0; d < WB; ++d) { const int x = lo + d; a[d] = (x >= 0 && x <= (int)L) ? KEY_ZERO : inf; }
			// codes of columns lo .. lo+WB-1 (0-based index lo-1 ..): NW unaligned groups of 8
			int cb = lo - 1;                                 // 0-based column index of the window's first nibble
			const uint32_t sh = (uint32_t)(cb & 7) * 4;
			int wi = cb >> 3;                                // arithmetic shift: floor for negative columns too
			uint32_t w0 = refw(wi);
			#pragma unroll
			for (int j = 0; j < NW; ++j) { const uint32_t w1 = refw(++wi); win[j] = __funnelshift_r(w0, w1, sh); w0 = w1; }
			// the codes that enter the window in rows 8g+1 .. 8g+8 are nibbles cb+WB+8g .. : the word pair (w0, w1) shifted by sh2
			constexpr int EXTRA = (NW * 8 - WB);             // nibbles of the last window word beyond the band: they are the first to enter
			const uint32_t sh2 = (uint32_t)((cb + WB) & 7) * 4;
			if (EXTRA) { wi = (cb + WB) >> 3; w0 = refw(wi); }
			win[NW - 1] &= TOPMASK;
			const int wbase = wi;                            // group g (rows 8g+1 .. 8g+8) needs the reference words wbase+g, wbase+g+1 and the query word g
			const uint32_t ngroups = (m + 7) >> 3;
			uint32_t bpre = k;                               // the slot's running minimum, fetched one group (8 rows) before it is applied
			// Fast rows: query and reference codes all plain bases, band inside the matrix, short query.  Then the substitution cost is
			// "the nibbles differ" (one XOR per row, no table), and cells above the budget need no clamp: they can never win or tie a
			// cell within it, and with m + WB < 480 no field of the key can overflow.
			const bool fastok = WMAX <= 32 && plain && m + WB + 8 < 480;
			int badrows = 0;                                 // rows during which the window may still hold a code that is not a plain base
			#pragma unroll
			for (int j = 0; j < NW; ++j) if (nonplain_nibbles(win[j] | (j == NW - 1 ? ~TOPMASK & 0x11111111u : 0u))) badrows = WB;
#ifndef EXT_NO_PREFETCH
			uint32_t w1n = refw(wbase + 1), qwn = qwd(0);
#endif
			for (uint32_t gq = 0; gq < ngroups && !dead; ++gq) {
				// tighten Emac as better hits land (burst.c:4159, 4220): a value read 8 rows ago is only less tight, never wrong
				k = min(k, bpre); inf = (k + 1) << 22;
				if (A.mode == BG_MODE_MIN && (gq & 3) == 0) bpre = __ldcg(A.best + slot);
#ifndef EXT_NO_PREFETCH
				const uint32_t w1 = w1n, qw = qwn;
				if (gq + 1 < ngroups) { w1n = refw(wbase + (int)gq + 2); qwn = qwd(gq + 1); }   // next group's words: their latency hides behind this group's rows
				const uint32_t feed = __funnelshift_r(w0, w1, sh2);
#else
				const uint32_t w1 = refw(wbase + (int)gq + 1);
				const uint32_t feed = __funnelshift_r(w0, w1, sh2), qw = qwd(gq);
#endif
				w0 = w1;
				const int x0g = (int)y + lo;                 // column (1-based) of band cell 0 in the first row of the group
				if (nonplain_nibbles(feed)) badrows = WB + 8;
				const bool fast = fastok && badrows == 0 && y + 7 <= m && x0g >= 1 && x0g + 7 + WB - 1 <= (int)L;
				badrows = max(0, badrows - 8);
				if (fast) {
					#pragma unroll
					for (int r = 0; r < 8; ++r) {
						const uint32_t qb = ((qw >> (4 * r)) & 15) * 0x11111111u;
						const uint32_t nc = (feed >> (4 * r)) & 15;
						#pragma unroll
						for (int j = 0; j < NW - 1; ++j) win[j] = __funnelshift_r(win[j], win[j + 1], 4);
						win[NW - 1] = (win[NW - 1] >> 4) | (nc << (4 * ((WB - 1) & 7)));
						uint32_t left = inf;
						#pragma unroll
						for (int j = 0; j < NW; ++j) {
							const uint32_t x = win[j] ^ qb;
							const uint32_t nz = (((x & 0x77777777u) + 0x77777777u) | x) & 0x88888888u, nzh = nz >> 16;   // bit 4d+3: cell d mismatches
							#pragma unroll
							for (int dd = 0; dd < 8; ++dd) {
								const int d = j * 8 + dd;
								if (d >= WB) break;
								const uint32_t st = dd < 5 ? (nz & (8u << (4 * dd))) * (1u << (19 - 4 * dd)) : (nzh & (8u << (4 * (dd - 4)))) * (1u << (19 - 4 * (dd - 4)));
								const uint32_t up = d + 1 < WB ? a[d + 1] : inf;
								uint32_t t = viaddmin(up, KEY_UP, a[d] + st);
								t = viaddmin(left, KEY_LEFT, t) & KEY_CLEAR;
								a[d] = t; left = t;
							}
						}
					}
					uint32_t rowmin = KEY_NONE;
					#pragma unroll
					for (int d = 0; d < WB; ++d) rowmin = min(rowmin, a[d]);
					y += 8;
					if (rowmin >= inf) { dead = true; y -= 1; }
				} else {
				#pragma unroll
				for (int r = 0; r < 8; ++r) {
					if (y > m) break;
					const int x0 = (int)y + lo;              // column (1-based) of band cell 0 in row y
					const uint32_t *Srow = sS + ((qw >> (4 * r)) & 15) * 16;
					uint32_t rowmin = KEY_NONE, left = inf;
					// slide the code window by one column
					const uint32_t nc = (feed >> (4 * r)) & 15;
					#pragma unroll
					for (int j = 0; j < NW - 1; ++j) win[j] = __funnelshift_r(win[j], win[j + 1], 4);
					win[NW - 1] = (win[NW - 1] >> 4) | (nc << (4 * ((WB - 1) & 7)));
					if (x0 >= 1 && x0 + WB - 1 <= (int)L) {  // interior row: every cell and predecessor is inside the matrix
						#pragma unroll
						for (int d = 0; d < WB; ++d) {
							const uint32_t st = Srow[(win[d >> 3] >> (4 * (d & 7))) & 15];
							const uint32_t up = d + 1 < WB ? a[d + 1] : inf;
							const uint32_t v = cell(a[d], up, left, st, inf);
							a[d] = v; left = v; rowmin = min(rowmin, v);
						}
					} else {
						#pragma unroll
						for (int d = 0; d < WB; ++d) {
							const int x = x0 + d;
							const uint32_t st = Srow[(win[d >> 3] >> (4 * (d & 7))) & 15];
							const uint32_t up = d + 1 < WB ? a[d + 1] : inf;
							uint32_t v = cell(a[d], up, left, st, inf);
							if (x < 0 || x > (int)L) v = inf;
							else if (x == 0) v = y <= k ? key_col0(y) : inf;
							a[d] = v; left = v; rowmin = min(rowmin, v);
						}
					}
					if (rowmin >= inf) { dead = true; break; }   // every lane cell > maxED: the reference truncates (burst.c:1062-1065)
					++y;
				}
				}
			}
			if (!dead) y = m + 1;
		} else if (W > A.smem_w && m + 1 <= A.scratch_w) {
			// Bands too wide for shared memory (runs of N in a genome make a prefix match for thousands of columns): swept in STRIPS of 32
			// matrix columns, left to right.  In matrix coordinates a cell needs (y-1, x-1), (y-1, x) and (y, x-1) -- its own strip or the
			// strip to the left -- so a strip keeps the previous row of its 32 columns in registers and only the strip's last column goes
			// through global memory (one value per row, read back by the next strip): register speed instead of two scratch accesses per
			// cell.  Cells outside the band or the matrix are "absent", exactly as the band sweep treats them; the last row is scanned as the
			// strips pass it, left to right.
			striped = true;
			uint32_t c_, q0_, n_;
			get_run(A.W, (sv.task >> 4) - A.W.run_base, c_, q0_, n_);
			const uint8_t *qs = A.codes + A.qi[q0_ + (sv.task & 15)].off;
			const int gnw = (int)((L + 7) >> 3);
			for (uint32_t yy = 0; yy <= m; ++yy) g[(size_t)yy * GT] = inf;                 // column "left of the first strip"
			const int xlast = min((int)m + lo + (int)W - 1, (int)L);
			for (int X0 = lo & ~31; X0 <= xlast; X0 += 32) {
				if (X0 + 31 < 0) continue;                                                  // left of the matrix: nothing but absent cells
				int y0 = X0 - lo - (int)W + 1; if (y0 < 0) y0 = 0;
				int y1 = X0 + 31 - lo; if (y1 > (int)m) y1 = (int)m;
				if (y1 < y0) continue;
				if (A.mode == BG_MODE_MIN) { k = min(k, __ldcg(A.best + slot)); inf = (k + 1) << 22; }
				uint32_t rw[4];                                                             // reference codes of matrix columns X0 .. X0+31 = lane nibbles X0-1 .. X0+30
				{ const int wi = (X0 - 1) >> 3; uint32_t w0 = lane_word_or0(lanew, wi, gnw);
				  #pragma unroll
				  for (int t = 0; t < 4; ++t) { const uint32_t w1 = lane_word_or0(lanew, wi + 1 + t, gnw); rw[t] = __funnelshift_r(w0, w1, 28); w0 = w1; } }
				uint32_t prev[32];
				uint32_t yy = (uint32_t)y0, bl_prev;
				if (y0 == 0) {                                                              // row 0 of the matrix (burst.c:723-725)
					#pragma unroll
					for (int j = 0; j < 32; ++j) { const int x = X0 + j; prev[j] = (x >= lo && x <= lo + (int)W - 1 && x >= 0 && x <= (int)L) ? KEY_ZERO : inf; }
					bl_prev = g[0]; g[0] = prev[31]; yy = 1;
				} else {
					#pragma unroll
					for (int j = 0; j < 32; ++j) prev[j] = inf;                             // row y0-1 ends left of this strip
					bl_prev = g[(size_t)(y0 - 1) * GT];
				}
				for (; yy <= (uint32_t)y1; ++yy) {
					const uint32_t *Srow = sS + (qs[yy - 1] & 15) * 16;
					const uint32_t bl_cur = g[(size_t)yy * GT];                             // (yy, X0-1), left there by the previous strip
					const int blo = (int)yy + lo, bhi = blo + (int)W - 1;                   // this row's band
					uint32_t left = bl_cur, diag = bl_prev;
					#pragma unroll
					for (int j = 0; j < 32; ++j) {
						const int x = X0 + j;
						const uint32_t up = prev[j];
						uint32_t v = cell(diag, up, left, Srow[(rw[j >> 3] >> (4 * (j & 7))) & 15u], inf);
						if (x < blo || x > bhi || x < 0 || x > (int)L) v = inf;
						else if (x == 0) v = yy <= k ? key_col0(yy) : inf;
						diag = up; prev[j] = v; left = v;
						if (yy == m && x >= blo && x <= bhi && x >= 1 && x <= (int)L) {     // last-row selection (burst.c:826-842, 863-883)
							const uint32_t kk = v >> 11;
							if (kk < sbk) { sbk = kk; sbshr = v & 0x1FF; sfp = (uint32_t)x; }
							else if (kk == sbk) sfp = (uint32_t)x;
						}
					}
					bl_prev = bl_cur; g[(size_t)yy * GT] = prev[31];
				}
				cells += 32ull * (unsigned)(y1 - y0 + 1);
			}
			y = m + 1;
		} else {
			uint32_t c_, q0_, n_;
			get_run(A.W, (sv.task >> 4) - A.W.run_base, c_, q0_, n_);
			const uint8_t *qs = A.codes + A.qi[q0_ + (sv.task & 15)].off;
			for (int d = 0; d < Wd; ++d) { const int x = lo + d; g[d * GT] = (x >= 0 && x <= (int)L) ? KEY_ZERO : inf; }
			const int gnwords = (int)((L + 7) >> 3); uint32_t cw = 0; int cwi = INT32_MIN;
			for (; y <= m; ++y) {
				const int x0 = (int)y + lo;
				const uint32_t *Srow = sS + (qs[y - 1] & 15) * 16;
				uint32_t rowmin = KEY_NONE, left = inf;
				uint32_t diag = g[0], upn = Wd > 1 ? g[GT] : inf;
				uint32_t cnext = (A.gen_mode & 1u) ? fetch_code(lanew, (uint32_t)(x0 - 1), L) : 0u;
				for (int d = 0; d < Wd; ++d) {
					const int x = x0 + d;
					const uint32_t up = upn;
					upn = d + 2 < Wd ? g[(d + 2) * GT] : inf;               // (read before this cell's store: cell d + 2 still holds the previous row)
					// reference code of column x: one 32-bit word of the lane serves 8 cells (the band slides one column per row,
					// so a load per cell would fetch every word W times over from L2)
					uint32_t code;
					if (A.gen_mode & 1u) { code = cnext; cnext = fetch_code(lanew, (uint32_t)x, L); }   // a load per cell, one cell ahead (default)
					else {
						const int col = x - 1, cwi_ = col >> 3;
						if (cwi_ != cwi) { cwi = cwi_; cw = lane_word_or0(lanew, cwi_, gnwords); }
						code = (col >= 0 && col < (int)L) ? (cw >> ((col & 7) * 4)) & 15u : 0u;
					}
					const uint32_t st = Srow[code];
					uint32_t v = cell(diag, up, left, st, inf);
					if (x < 0 || x > (int)L) v = inf;
					else if (x == 0) v = y <= k ? key_col0(y) : inf;
					diag = up; g[d * GT] = v; left = v; rowmin = min(rowmin, v);
				}
				if (rowmin >= inf) { dead = true; break; }
				if ((y & 15) == 0 && A.mode == BG_MODE_MIN) { k = min(k, A.best[slot]); inf = (k + 1) << 22; }
			}
		}
		if (!striped) cells += (unsigned long long)(dead ? y : m) * (unsigned)Wd;
		uint32_t out = 0, fp = sfp;
		if (!dead) {
			// last-row selection, left to right (burst.c:826-842, 863-883)
			uint32_t bk = sbk, bshr = sbshr;
			auto scan = [&](int d, uint32_t v) {
				const int x = (int)m + lo + d;
				if (x < 1 || x > (int)L) return;
				const uint32_t kk = v >> 11;
				if (kk < bk) { bk = kk; bshr = v & 0x1FF; fp = (uint32_t)x; }
				else if (kk == bk) fp = (uint32_t)x;
			};
			if (WMAX) {
				#pragma unroll
				for (int d = 0; d < WB; ++d) if (d < (int)W) scan(d, a[d]);      // the cluster's own diagonals only: beyond them the sweep may see part of a neighbouring cluster's band
			} else if (!striped) for (int d = 0; d < Wd; ++d) scan(d, g[d * GT]);
			const uint32_t ed = bk >> 11, sh = 2047u - (bk & 2047u);
			if (bk != (KEY_NONE >> 11) && ed <= k) {
				out = ed | (sh << 8) | (bshr << 16) | (1u << 31);
				atomicMin(&A.best[slot], ed);
			}
		}
		Res r; r.a = out; r.b = fp; r.slot = slot;
		A.res[i] = r;
	}
	if (cells) atomicAdd(A.band_cells, cells);
}

// ---------------------------------------------------------------------------------------------
// Phase C: keep what the reference keeps.  The clusters of one (task, lane) cover disjoint diagonal
// ranges, hence disjoint stretches of the last row, in ascending column order: the reference's
// left-to-right scan (burst.c:826-883) keeps the best (score, shift), numGapR of its first
// occurrence and the column of its last.
// ---------------------------------------------------------------------------------------------
__global__ void k_select(const Surv *__restrict__ surv, const Res *__restrict__ res,
		const uint32_t *__restrict__ best, uint32_t *counters, uint32_t surv_cap, bg_hit *__restrict__ hits,
		unsigned long long *__restrict__ keys, int mode) {
	const uint32_t nsurv = min(counters[C_SURV], surv_cap);
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nsurv; i += gridDim.x * blockDim.x) {
		const Surv sv = surv[i];
		const uint32_t grp = (sv.w_lane >> 4) & 15;
		if (!grp) continue;
		uint32_t bkey = 0xFFFFFFFFu, gr = 0, fp = 0;
		for (uint32_t s = 0; s < grp && i + s < nsurv; ++s) {
			const Res r = res[i + s];
			if (!(r.a >> 31)) continue;
			const uint32_t key = ((r.a & 255) << 8) | (255 - ((r.a >> 8) & 255));
			if (key < bkey) { bkey = key; gr = (r.a >> 16) & 255; fp = r.b; }
			else if (key == bkey) fp = r.b;
		}
		if (bkey == 0xFFFFFFFFu) continue;
		const uint32_t ed = bkey >> 8;
		if (mode == BG_MODE_MIN && ed != best[res[i].slot]) continue;     // burst.c:4229, 4497
		const uint32_t j = atomicAdd(&counters[C_HITS], 1u);
		bg_hit h; h.task = sv.task; h.lane = (uint8_t)(sv.w_lane & 15); h.ed = (uint8_t)ed;
		h.gap_q = (uint8_t)(255 - (bkey & 255)); h.gap_r = (uint8_t)gr; h.final_pos = fp;
		hits[j] = h;
		keys[j] = ((unsigned long long)sv.task << 4) | (sv.w_lane & 15);
	}
}

__global__ void k_gather_hits(const bg_hit *__restrict__ in, const uint32_t *__restrict__ order, uint32_t n, bg_hit *__restrict__ out) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = in[order[i]];
}
__global__ void k_iota(uint32_t *v, uint32_t n) { uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) v[i] = i; }

__global__ void k_init_best(uint32_t *best, const uint16_t *in, uint32_t n) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) best[i] = in ? in[i] : 0xFFFFu;
}

// ---------------------------------------------------------------------------------------------
// Compact strand batches (bg_align_bunches_into): every READ crosses the bus once, 2 or 4 bits per base; the device
// derives both strands (burst.c:3087-3109 builds the reverse-complement copies on the host), the per-strand records
// and the run list from the bunch -> candidate lists the reference's driver works with (burst.c:4085-4157).
// ---------------------------------------------------------------------------------------------
__constant__ uint8_t c_rvt[16] = {0, 4, 3, 2, 1, 5, 7, 6, 9, 8, 10, 11, 13, 12, 15, 14};       // burst.c:168
// lengths: per read (u16 -> u64) and per strand, the latter rounded up to 16 so that every strand's codes start 16-byte aligned
__global__ void k_compact_rlen(const uint16_t *__restrict__ rlen, uint32_t nreads, unsigned long long *__restrict__ rl64) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i <= nreads) rl64[i] = i < nreads ? rlen[i] : 0;
}
__global__ void k_compact_slen(const uint16_t *__restrict__ rlen, uint32_t nreads, const uint32_t *__restrict__ strand, uint32_t nq,
		unsigned long long *__restrict__ sl64, uint32_t *__restrict__ counters) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > nq) return;
	unsigned long long v = 0;
	if (i < nq) {
		const uint32_t r = strand[i] & 0x7FFFFFFFu;
		if (r >= nreads) atomicExch(&counters[C_ERR], i + 1);
		else v = ((unsigned long long)rlen[r] + 15ull) & ~15ull;
	}
	sl64[i] = v;
}
// codes of every strand (one byte per base, 16-byte aligned start): 8 threads per strand, 16 bases per thread and step
__global__ void k_compact_codes(const uint8_t *__restrict__ reads, uint32_t flags, const unsigned long long *__restrict__ roff, const uint16_t *__restrict__ rlen,
		const uint32_t *__restrict__ strand, const unsigned long long *__restrict__ qoff, uint32_t nq, uint32_t nreads, uint8_t *__restrict__ codes) {
	const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, sub = threadIdx.x & 7;
	if (q >= nq) return;
	const uint32_t sv = strand[q], r = sv & 0x7FFFFFFFu; const bool rc = sv >> 31;
	if (r >= nreads) return;
	const uint32_t len = rlen[r];
	const unsigned long long ro = roff[r], qo = qoff[q];
	const bool two = (flags & BG_R_PACKED2) != 0;
	for (uint32_t b0 = sub * 16; b0 < len; b0 += 128) {
		uint32_t w[4] = {0, 0, 0, 0};
		#pragma unroll
		for (int i = 0; i < 16; ++i) {
			const uint32_t pos = b0 + i;
			if (pos < len) {
				const unsigned long long x = ro + (rc ? len - 1 - pos : pos);
				uint32_t code = two ? ((__ldg(reads + (x >> 2)) >> (2 * (x & 3))) & 3u) + 1u : (__ldg(reads + (x >> 1)) >> (4 * (x & 1))) & 15u;
				if (rc) code = c_rvt[code];
				w[i >> 2] |= code << (8 * (i & 3));
			}
		}
		*(uint4 *)(codes + qo + b0) = make_uint4(w[0], w[1], w[2], w[3]);
	}
}
// The two kernels above in one pass for reads that arrive at 2 bits per base (every base a plain A/C/G/T): 8 threads per strand, 16 bases per
// thread and step, read as ONE unaligned 32-bit window of the packed stream (two aligned loads + a funnel shift) instead of 16 byte loads;
// the reverse complement is the window of the mirrored position with its 2-bit fields reversed (brev + swap within fields) and inverted
// (A<->T, C<->G = 3 - code).  The thread writes its 16 code bytes (one 128-bit store) and its two words of 8 nibbles, the strand's first
// thread the seed class and the counters k_qprep keeps.  0.65 ms -> 0.08 ms per 2 M strands of 100 bases.
__device__ __forceinline__ unsigned long long spread2_to_bytes(uint32_t v16) {      // 8 fields of 2 bits -> 8 bytes
	unsigned long long x = v16 & 0xFFFFu;
	x = (x | (x << 24)) & 0x000000FF000000FFull;
	x = (x | (x << 12)) & 0x000F000F000F000Full;
	x = (x | (x << 6)) & 0x0303030303030303ull;
	return x;
}
__device__ __forceinline__ uint32_t spread2_to_nibbles(uint32_t v16) {               // 8 fields of 2 bits -> 8 nibbles
	uint32_t x = v16 & 0xFFFFu;
	x = (x | (x << 8)) & 0x00FF00FFu;
	x = (x | (x << 4)) & 0x0F0F0F0Fu;
	x = (x | (x << 2)) & 0x33333333u;
	return x;
}
__global__ void __launch_bounds__(256) k_compact_prep2(const uint32_t *__restrict__ reads32, const unsigned long long *__restrict__ roff, const uint16_t *__restrict__ rlen,
		const uint16_t *__restrict__ rbudget, const uint32_t *__restrict__ strand, const unsigned long long *__restrict__ qoff, uint32_t nq, uint32_t nreads,
		uint8_t *__restrict__ codes, QInfo *__restrict__ qi, SeedLayout SL, uint32_t *__restrict__ qnib, uint32_t *__restrict__ nseed, uint32_t *__restrict__ unseeded) {
	// a resident grid walks the strands; the three counters are summed per block in shared memory and leave through ONE set of atomics per block
	// (one set per warp -- 1.5 M atomics on three addresses for 2 M strands -- took longer than everything else in the kernel: 1.1 ms)
	__shared__ uint32_t sh[4];
	if (threadIdx.x < 4) sh[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t sub = threadIdx.x & 7;
	for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; q < ((nq + 31u) & ~31u); q += (gridDim.x * blockDim.x) >> 3) {
	bool ok = q < nq, seed = false; uint32_t nst = 0;
	uint32_t r = 0; bool rc = false;
	if (ok) { const uint32_t sv = strand[q]; r = sv & 0x7FFFFFFFu; rc = sv >> 31; ok = r < nreads; }
	if (ok) {
		const uint32_t len = rlen[r];
		const unsigned long long ro = roff[r], qo = qoff[q];
		uint32_t *W = qnib + (qo >> 3) + 3ull * q;
		if (sub == 0) { W[0] = 0; W[1] = 0; }
		for (uint32_t b0 = sub * 16; b0 < len; b0 += 128) {
			const uint32_t n = min(16u, len - b0);
			uint32_t v;
			if (!rc) {
				const unsigned long long bit = 2ull * (ro + b0);
				const uint32_t *pw = reads32 + (bit >> 5);
				v = __funnelshift_r(__ldg(pw), __ldg(pw + 1), (uint32_t)bit & 31u);     // (the stream is padded: the second word exists)
			} else {
				const long long xs = (long long)ro + (long long)len - 16 - (long long)b0;   // first base of the mirrored window; field f = base xs + f
				if (xs >= 0) {
					const unsigned long long bit = 2ull * (unsigned long long)xs;
					const uint32_t *pw = reads32 + (bit >> 5);
					v = __funnelshift_r(__ldg(pw), __ldg(pw + 1), (uint32_t)bit & 31u);
				} else v = __ldg(reads32) << (2 * (uint32_t)(-xs));                         // (only the first read of the stream: fields below its first base are not used)
				v = __brev(v);
				v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
				v = ~v;
			}
			// code bytes (1..4), zero beyond the strand's end, as k_compact_codes leaves them
			unsigned long long lo = spread2_to_bytes(v) + 0x0101010101010101ull, hi = spread2_to_bytes(v >> 16) + 0x0101010101010101ull;
			if (n < 16) {
				if (n <= 8) { hi = 0; if (n < 8) lo &= (1ull << (8 * n)) - 1ull; }
				else hi &= (1ull << (8 * (n - 8))) - 1ull;
			}
			*(uint4 *)(codes + qo + b0) = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
			// the same bases as nibbles, the last word cut at the strand's end (k_qprep)
			uint32_t w0 = spread2_to_nibbles(v) + 0x11111111u, w1 = spread2_to_nibbles(v >> 16) + 0x11111111u;
			if (n < 8) w0 &= (1u << (4 * n)) - 1u;
			else if (n < 16 && n > 8) w1 &= (1u << (4 * (n - 8))) - 1u;
			W[2 + (b0 >> 3)] = w0;
			if (n > 8) W[3 + (b0 >> 3)] = w1;
		}
		if (sub == 0) {
			const uint32_t k = rbudget[r], np = k + 1u, plen = len / np;
			seed = SL.stride && np <= SL.np_max && plen >= SL.w + SL.stride - 1;           // every base is plain: nothing else to check
			qi[q].cls = (uint8_t)((seed ? 1u : 0u) | 2u);
			if (seed) nst = np;
		}
	}
	const uint32_t m = __ballot_sync(0xFFFFFFFFu, seed), tot = __reduce_add_sync(0xFFFFFFFFu, nst), mx = __reduce_max_sync(0xFFFFFFFFu, nst);
	if (m && (threadIdx.x & 31) == 0) { atomicAdd(&sh[0], (uint32_t)__popc(m)); atomicAdd(&sh[1], tot); atomicMax(&sh[2], mx); }
	const uint32_t u = __ballot_sync(0xFFFFFFFFu, q < nq && sub == 0 && !seed);
	if (u && (threadIdx.x & 31) == 0) atomicAdd(&sh[3], (uint32_t)__popc(u));
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		if (sh[0]) { atomicAdd(nseed, sh[0]); atomicAdd(nseed + 1, sh[1]); atomicMax(nseed + 2, sh[2]); }
		if (unseeded && sh[3]) atomicAdd(unseeded, sh[3]);
	}
}
// the strand records (as k_qinfo) and the histogram of stretch lengths
__global__ void k_compact_qinfo(const unsigned long long *__restrict__ qoff, const uint16_t *__restrict__ rlen, const uint16_t *__restrict__ rbudget,
		const uint32_t *__restrict__ strand, uint32_t nq, uint32_t nreads, QInfo *__restrict__ qi, uint32_t *__restrict__ hist, uint32_t *__restrict__ counters) {
	__shared__ uint32_t sh[32];
	if (threadIdx.x < 32) sh[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q < nq) {
		uint32_t r = strand[q] & 0x7FFFFFFFu, len = 1, k = 0;
		if (r >= nreads) { atomicExch(&counters[C_ERR], q + 1); r = 0; }
		else { len = rlen[r]; k = rbudget[r]; if (!len || k > 254) { atomicExch(&counters[C_ERR], q + 1); len = 1; k = 0; } }
		QInfo Q; Q.off = qoff[q]; Q.len = len; Q.slot = r; Q.k = (uint16_t)k; Q.P = filt_code(len, k); Q.cls = 0;
		qi[q] = Q;
		if (k + 1 <= SEED_NP_MAX) atomicAdd(&sh[min(len / (k + 1), 31u)], 1u);
	}
	__syncthreads();
	if (threadIdx.x < 32 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
// bunch -> candidate lists into runs: run r = candidate r of the bunch that owns it (cand_off), all queries of the bunch; runs r0 .. r0+count-1
__global__ void k_compact_runs(const uint32_t *__restrict__ cand_off, const uint32_t *__restrict__ cand, uint32_t nbunch, uint32_t r0, uint32_t count, uint32_t qbunch, uint32_t nq,
		bg_run *__restrict__ runs, uint32_t *__restrict__ counters) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const uint32_t r = r0 + i;
	uint32_t lo = 0, hi = nbunch;                                          // last bunch b with cand_off[b] <= r
	while (lo + 1 < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(cand_off + mid) <= r) lo = mid; else hi = mid; }
	const unsigned long long q0 = (unsigned long long)lo * qbunch;
	bg_run R; R.clump = cand[r]; R.query0 = (uint32_t)min(q0, (unsigned long long)nq); R.nq = q0 < nq ? (uint32_t)min((unsigned long long)qbunch, nq - q0) : 0u;
	if (!R.nq || __ldg(cand_off + lo) > r || __ldg(cand_off + lo + 1) <= r) { atomicExch(&counters[C_ERR], 0x80000000u | r); R.nq = 1; R.query0 = min(R.query0, nq - 1); }
	runs[i] = R;
}

// ---------------------------------------------------------------------------------------------
// Candidate generation on the device (burst.c:4085-4133 with postScour20/24, 3238-3282).
// One block per bunch of <= 16 strands, in the reference's steps:
//   words    every N-mer of every query -> (word, query) pairs                              burst.c:4097-4105
//   sort     by (word, query)                                                               burst.c:4118
//   count    each distinct word adds its largest per-query multiplicity to every clump of its posting list; the clumps touched
//            are remembered in first-touch order (words ascending, postings in list order)   burst.c:3238-3282
//   pick     clumps whose count exceeds the bunch's smallest threshold len - (ed+1) N, ordered by descending count (stable: ties keep
//            first-touch order, as iSort / qsort do)                                        burst.c:4120-4130, 4038-4046
//   emit     one run per candidate and maximal range of queries whose own threshold it passes, then the always-visited BadList
//                                                                                            burst.c:4137-4168, 4281-4283
// The per-clump counters are a dense array in global memory per resident block (the reference's Hash[] / Cache[] per thread), cleared
// through the touched list.  Queries must be plain A/C/G/T (the reference expands ambiguous bases into every variant, burst.c:4106-4113;
// such batches stay on the host path).
// ---------------------------------------------------------------------------------------------
struct CandArgs {
	const QInfo *qi; const uint32_t *qnib; uint32_t nq, qbunch, nbunch;
	const unsigned long long *acx_off; const uint8_t *post; int big, N; uint32_t num_clumps;
	const uint32_t *bad; uint32_t nbad; int heur, skip_bad;
	uint32_t *cnt, *first, *cache;              // per resident block: num_clumps counters, first-touch keys, touched list
	bg_run *runs; uint32_t runs_cap; uint32_t *counters;   // counters[C_RUNS] = runs emitted (keeps counting past the capacity)
	uint32_t pmax, cmax;                         // capacity of the pair array / candidate list in shared memory (powers of two)
};
enum { C_RUNS = 5 };

__device__ __forceinline__ void bitonic_sort_u64(unsigned long long *a, uint32_t n) {   // n a power of two, whole block
	for (uint32_t k = 2; k <= n; k <<= 1)
		for (uint32_t j = k >> 1; j > 0; j >>= 1) {
			for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
				const uint32_t l = i ^ j;
				if (l > i) {
					const unsigned long long x = a[i], y = a[l];
					if (((i & k) == 0) == (x > y)) { a[i] = y; a[l] = x; }
				}
			}
			__syncthreads();
		}
}
__device__ __forceinline__ uint32_t posting_at(const uint8_t *p, uint32_t e, int big) {
	if (big) { const uint8_t *q = p + (size_t)e * 3; return (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16); }
	const uint8_t *q = p + (size_t)(e >> 1) * 5;
	if (!(e & 1)) return ((uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16)) & 0xFFFFFu;
	return (((uint32_t)q[2] >> 4) | ((uint32_t)q[3] << 4) | ((uint32_t)q[4] << 12)) & 0xFFFFFu;
}

__global__ void __launch_bounds__(128) k_candgen(CandArgs A) {
	extern __shared__ __align__(16) unsigned long long csm[];
	unsigned long long *pairs = csm;                       // pmax: (word << 5 | query), later reused for the candidate keys
	__shared__ uint32_t s_len[16], s_mm[16], s_off[17], s_np, s_ncache, s_ncand, s_minmm, s_base, s_nq, s_nlong, s_tot;
	__shared__ uint32_t s_long[128];                        // heads whose posting lists are long: walked by whole warps
	const uint32_t N = (uint32_t)A.N;
	uint32_t *cnt = A.cnt + (size_t)blockIdx.x * A.num_clumps, *first = A.first + (size_t)blockIdx.x * A.num_clumps, *cache = A.cache + (size_t)blockIdx.x * A.num_clumps;
	uint32_t *candp = (uint32_t *)(csm + A.pmax);          // cmax clump ids, parallel to the candidate keys
	for (uint32_t b = blockIdx.x; b < A.nbunch; b += gridDim.x) {
		const uint32_t z = b * A.qbunch, nb = min(A.qbunch, A.nq - z);
		if (threadIdx.x == 0) { s_np = 0; s_ncache = 0; s_ncand = 0; s_minmm = 0xFFFFFFFFu; s_nq = nb; s_nlong = 0; }
		__syncthreads();
		if (threadIdx.x < nb) {                                // thresholds (burst.c:4091-4095, 4163-4164)
			const QInfo Q = A.qi[z + threadIdx.x];
			const uint32_t len = Q.len, kload = (uint32_t)Q.k * N + N;
			uint32_t mmatch = kload < len ? len - kload : 0;
			const uint32_t heur = A.heur ? (len >> 4) + 1u : 0u;
			if (mmatch < heur) mmatch = heur;
			atomicMin(&s_minmm, mmatch);
			s_mm[threadIdx.x] = kload < len ? len - kload : 1;
			s_len[threadIdx.x] = len;
		}
		__syncthreads();
		// ---- words ----
		if (threadIdx.x == 0) {
			uint32_t t = 0;
			for (uint32_t j = 0; j < nb; ++j) { s_off[j] = t; t += s_len[j] >= N ? s_len[j] - N + 1 : 0; }
			s_off[nb] = t;
			if (t > A.pmax) { atomicExch(&A.counters[C_ERR], 0x20000000u | b); t = 0; for (uint32_t j = 0; j <= nb; ++j) s_off[j] = 0; }   // queries longer than the pair array was sized for
			s_np = t;
		}
		__syncthreads();
		for (uint32_t j = 0; j < nb; ++j) {
			const QInfo Q = A.qi[z + j];
			const uint32_t *Wq = A.qnib + (Q.off >> 3) + 3ull * (z + j) + 2;
			const uint32_t base = s_off[j], nw = s_off[j + 1] - base;
			for (uint32_t p = threadIdx.x; p < nw; p += blockDim.x) {
				unsigned long long w = 0;
				for (uint32_t t = 0; t < N; ++t) { const uint32_t x = p + t; w = (w << 2) | (((Wq[x >> 3] >> (4 * (x & 7))) & 15u) - 1u); }
				pairs[base + p] = (w << 5) | j;
			}
		}
		__syncthreads();
		const uint32_t np = s_np;
		uint32_t n2 = 2; while (n2 < np) n2 <<= 1;
		for (uint32_t i = np + threadIdx.x; i < n2; i += blockDim.x) pairs[i] = ~0ull;
		__syncthreads();
		bitonic_sort_u64(pairs, n2);
		// ---- count: thread i owns the distinct word that starts at i ----
		for (uint32_t i0 = 0; i0 < np; i0 += blockDim.x) {
			const uint32_t i = i0 + threadIdx.x;
			if (i < np) {
				const unsigned long long w = pairs[i] >> 5;
				if (i == 0 || (pairs[i - 1] >> 5) != w) {
					uint32_t mx = 0, e = i;
					while (e < np && (pairs[e] >> 5) == w) { uint32_t r = e; while (r < np && pairs[r] == pairs[e]) ++r; mx = max(mx, r - e); e = r; }
					const unsigned long long o0 = A.acx_off[w], o1 = A.acx_off[w + 1];
					const uint32_t L = A.big ? (uint32_t)((o1 - o0) / 3) : (uint32_t)(((o1 - o0) / 5) * 2 + ((o1 - o0) % 5 ? 1 : 0));
					if (L > 64) { const uint32_t s = atomicAdd(&s_nlong, 1u); if (s < 128) s_long[s] = i | (mx << 16); else {
						const uint8_t *pp = A.post + o0;                  // (queue full: walk it alone)
						for (uint32_t t = 0; t < L; ++t) { const uint32_t c = posting_at(pp, t, A.big); if (c < A.num_clumps) { if (atomicAdd(&cnt[c], mx) == 0) cache[atomicAdd(&s_ncache, 1u)] = c; atomicMin(&first[c], (i << 19) | min(t, 0x7FFFFu)); } }
					} }
					else {
						const uint8_t *pp = A.post + o0;
						for (uint32_t t = 0; t < L; ++t) { const uint32_t c = posting_at(pp, t, A.big); if (c < A.num_clumps) { if (atomicAdd(&cnt[c], mx) == 0) cache[atomicAdd(&s_ncache, 1u)] = c; atomicMin(&first[c], (i << 19) | min(t, 0x7FFFFu)); } }
					}
				}
			}
			__syncthreads();
			// long posting lists of this round: a warp each
			const uint32_t nl = min(s_nlong, 128u);
			for (uint32_t s = threadIdx.x >> 5; s < nl; s += blockDim.x >> 5) {
				const uint32_t i = s_long[s] & 0xFFFFu, mx = s_long[s] >> 16;
				const unsigned long long w = pairs[i] >> 5, o0 = A.acx_off[w], o1 = A.acx_off[w + 1];
				const uint32_t L = A.big ? (uint32_t)((o1 - o0) / 3) : (uint32_t)(((o1 - o0) / 5) * 2 + ((o1 - o0) % 5 ? 1 : 0));
				const uint8_t *pp = A.post + o0;
				for (uint32_t t = threadIdx.x & 31; t < L; t += 32) { const uint32_t c = posting_at(pp, t, A.big); if (c < A.num_clumps) { if (atomicAdd(&cnt[c], mx) == 0) cache[atomicAdd(&s_ncache, 1u)] = c; atomicMin(&first[c], (i << 19) | min(t, 0x7FFFFu)); } }
			}
			__syncthreads();
			if (threadIdx.x == 0) s_nlong = 0;
			__syncthreads();
		}
		// ---- pick: candidates sorted by (count descending, first touch ascending); the key carries the slot of the clump id ----
		const uint32_t ncache = s_ncache, minmm = s_minmm;
		unsigned long long *ckey = pairs;                       // (the pairs are done with)
		__syncthreads();
		for (uint32_t i = threadIdx.x; i < ncache; i += blockDim.x) {
			const uint32_t c = cache[i], v = min(cnt[c], 65535u), f = first[c];
			cnt[c] = 0; first[c] = 0xFFFFFFFFu;
			if (v > minmm) {
				const uint32_t s = atomicAdd(&s_ncand, 1u);
				if (s < A.cmax) { candp[s] = c; ckey[s] = ((unsigned long long)(65535u - v) << 48) | ((unsigned long long)f << 16) | s; }
				else atomicExch(&A.counters[C_ERR], 0x40000000u | b);    // more candidates than the list holds: reported, the call fails (raise cmax)
			}
		}
		__syncthreads();
		const uint32_t ncand = min(s_ncand, A.cmax);
		uint32_t c2 = 2; while (c2 < ncand) c2 <<= 1;
		for (uint32_t i = ncand + threadIdx.x; i < c2; i += blockDim.x) ckey[i] = ~0ull;
		__syncthreads();
		if (ncand > 1) bitonic_sort_u64(ckey, c2);
		// ---- emit: per candidate the maximal ranges of queries whose own threshold it passes (burst.c:4163-4168), then the BadList ----
		uint32_t *nsr = candp + A.cmax;                          // sub-runs per candidate, then their exclusive prefix
		for (uint32_t s = threadIdx.x; s < ncand; s += blockDim.x) {
			const uint32_t v = 65535u - (uint32_t)(ckey[s] >> 48);
			uint32_t r = 0; bool in = false;
			for (uint32_t j = 0; j < nb; ++j) { const bool p = v > s_mm[j]; r += p && !in; in = p; }
			nsr[s] = r;
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			uint32_t t = 0;
			for (uint32_t s = 0; s < ncand; ++s) { const uint32_t r = nsr[s]; nsr[s] = t; t += r; }
			const uint32_t tot = t + (A.skip_bad ? 0u : A.nbad);
			s_tot = t;
			s_base = tot ? atomicAdd(&A.counters[C_RUNS], tot) : 0u;
		}
		__syncthreads();
		const uint32_t base = s_base;
		for (uint32_t s = threadIdx.x; s < ncand; s += blockDim.x) {
			const uint32_t v = 65535u - (uint32_t)(ckey[s] >> 48), clump = candp[(uint32_t)ckey[s] & 0xFFFFu];
			uint32_t o = base + nsr[s], a0 = 0;
			while (a0 < nb) {
				while (a0 < nb && !(v > s_mm[a0])) ++a0;
				uint32_t b0 = a0;
				while (b0 < nb && v > s_mm[b0]) ++b0;
				if (b0 > a0) { if (o < A.runs_cap) { bg_run R; R.clump = clump; R.query0 = z + a0; R.nq = b0 - a0; A.runs[o] = R; } ++o; }
				a0 = b0;
			}
		}
		if (!A.skip_bad) for (uint32_t s = threadIdx.x; s < A.nbad; s += blockDim.x) {
			const uint32_t o = base + s_tot + s;
			if (o < A.runs_cap) { bg_run R; R.clump = A.bad[s]; R.query0 = z; R.nq = nb; A.runs[o] = R; }           // (ids outside the loaded clump range are skipped by every kernel)
		}
		__syncthreads();
	}
}

// hits of a device-generated run list -> (strand, clump) form
__global__ void k_xhits(const bg_hit *__restrict__ in, const bg_run *__restrict__ runs, uint32_t n, bg_xhit *__restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const bg_hit h = in[i]; const bg_run R = runs[h.task >> 4];
	bg_xhit x; x.query = R.query0 + (h.task & 15); x.clump = R.clump; x.lane = h.lane; x.ed = h.ed; x.gap_q = h.gap_q; x.gap_r = h.gap_r; x.final_pos = h.final_pos;
	out[i] = x;
}
// bytes of every posting list from the on-disk lengths (burst.c:3504-3527)
__global__ void k_acx_sizes(const uint32_t *__restrict__ lens, unsigned long long nk, int big, unsigned long long *__restrict__ out) {
	const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i > nk) return;
	const uint32_t L = i < nk ? lens[i] : 0u;
	out[i] = big ? (unsigned long long)L * 3 : (unsigned long long)(L / 2u) * 5 + (L & 1u) * 3;
}

// run validation (explicit run lists): malformed runs raise the error flag
__global__ void k_check_runs(const bg_run *__restrict__ runs, uint64_t nruns, uint32_t q_base, uint32_t nq, uint32_t *counters) {
	uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nruns) return;
	const bg_run R = runs[r];
	if (!R.nq || R.nq > BG_RUN_MAX || R.query0 < q_base || (uint64_t)R.query0 - q_base + R.nq > nq) atomicExch(&counters[C_ERR], 0x80000000u | (uint32_t)min(r, (uint64_t)0x7FFFFFFF));
}

// work statistics: nominal = sum over tasks of 16 * qlen * ClumpLen
__global__ void k_work_stats(Work W, const QInfo *__restrict__ qi, const uint32_t *__restrict__ clump_len, unsigned long long *out) {
	unsigned long long tasks = 0, nominal = 0, fcells = 0, scells = 0;
	for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < W.nruns; r += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t c, q0, n;
		if (!get_run(W, r, c, q0, n)) continue;
		const unsigned long long L = clump_len[c];
		bool seeded = false;
		for (uint32_t i = 0; i < n; ++i) {
			const QInfo Q = qi[q0 + i];
			nominal += 16ull * Q.len * L;
			if (Q.cls & 1) seeded = true; else fcells += 16ull * filt_rows(Q.P) * L;
		}
		if (seeded) scells += 16ull * L;                         // k_seed streams the clump once per run
		tasks += n;
	}
	for (int o = 16; o; o >>= 1) {
		tasks += __shfl_down_sync(0xFFFFFFFFu, tasks, o); nominal += __shfl_down_sync(0xFFFFFFFFu, nominal, o);
		fcells += __shfl_down_sync(0xFFFFFFFFu, fcells, o); scells += __shfl_down_sync(0xFFFFFFFFu, scells, o);
	}
	if ((threadIdx.x & 31) == 0) { atomicAdd(out, tasks); atomicAdd(out + 1, nominal); atomicAdd(out + 2, fcells); atomicAdd(out + 3, scells); }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename T> struct DBuf {
	T *p = nullptr; size_t cap = 0; bool borrowed = false;       // borrowed: another context's allocation (bg_share_db), never freed or resized here
	int need(size_t n) {
		if (n <= cap && !borrowed) return 0;
		if (p && !borrowed) cudaFree(p);
		p = nullptr; cap = 0; borrowed = false;
		size_t want = n + n / 8 + 64;
		cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
		if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(BG_ENOMEM, "cudaMalloc(%zu bytes): %s", want * sizeof(T), cudaGetErrorString(e)); }
		cap = want; return 0;
	}
	void release() { if (p && !borrowed) cudaFree(p); p = nullptr; cap = 0; borrowed = false; }
	void borrow(const DBuf<T> &o) { release(); p = o.p; cap = o.cap; borrowed = o.p != nullptr; }
};

enum WorkKind { WORK_NONE = 0, WORK_ALL = 1, WORK_TASKS = 2, WORK_RUNS = 3 };

// One of the device buffer sets the pipelined one-call path cycles through: while the kernels of one slice run, the
// next slices' queries and runs are copied in on a second stream (three sets: the copies never wait for the kernels).
#define NSLICEBUF 3
struct Slice {
	DBuf<uint8_t> packed, codes; DBuf<uint64_t> qoff; DBuf<uint16_t> budget; DBuf<uint32_t> slot;
	DBuf<QInfo> qi; DBuf<uint32_t> qnib; DBuf<bg_run> runs; DBuf<unsigned long long> sl64;
	cudaEvent_t copied = nullptr, computed = nullptr;
	void release() { sl64.release(); packed.release(); codes.release(); qoff.release(); budget.release(); slot.release(); qi.release(); qnib.release(); runs.release(); }
};

// Contexts that share one database (bg_share_db) also share a gate: the kernel sequence of one batch is queued behind the kernel sequence of
// the batch queued before it, whichever context that was, so that the batches take the GPU one after the other while the copies of the next
// one travel.  Without it two host threads that start together stay in lockstep -- both copy, then both compute with half the SMs each, then
// both copy back -- and nothing overlaps (measured: 3.14 ms per step against 3.60 ms for one context).
struct Gate {
	std::mutex m; cudaEvent_t last = nullptr; std::vector<cudaEvent_t> ev;
	~Gate() { for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e); }
};

struct bg_ctx {
	int device = 0;
	std::shared_ptr<Gate> gate; cudaEvent_t gate_done = nullptr;
	cudaStream_t stream = nullptr; bool own_stream = false;
	int sms = 148;
	int seed_filter = 1, seed_chunk = 8, seed_words = 0, seed_stage = 0;   // tuning: runs per warp, Bloom words per warp (0 = auto), bulk-copy staging
	int seed_groups = 0;                                          // groups (runs per round) per block, 0 = chosen for occupancy
	int seed_impl = 1, seed_nch = 0, seed_nch_auto = 8, seed_lbits = 0, seed_fb = 2, seed_hslots = 0, seed_vmode = 0;   // seed_nch 0: seed_nch_auto, chosen from the database's clump lengths at load;   // seed_fb: filter bits per window (1 or 2)
	int _pad0 = 0;              // 1: warp-per-bunch k_seedw (default), 0: block form k_seed; chunks per register buffer; log2 bitmap bits (0 = from the batch)
	uint32_t mstage = 0;                                          // longest query the k_extend staging slots are sized for (from the batch's lengths)
	uint32_t seed_npmax = 1;                                      // stretches per query the window table holds (from the batch)
	bool seed_ok = true; uint32_t amb_add = 0x22222222u, m16[8];   // derived from the scoring table
	// scoring
	uint8_t S[256];
	DBuf<uint32_t> d_sterm;
	// DB
	DBuf<uint4> d_db; DBuf<uint64_t> d_clump_off; DBuf<uint32_t> d_clump_len; DBuf<ClumpMeta> d_meta;
	uint32_t stage_bytes = 0;                                     // k_seed staging buffer: the largest clump, at most 8 KB
	uint32_t num_clumps = 0, first_clump = 0;
	// batch
	DBuf<uint8_t> d_packed; DBuf<uint8_t> d_codes; DBuf<uint64_t> d_qoff; DBuf<uint16_t> d_budget; DBuf<uint32_t> d_slot;
	DBuf<QInfo> d_qi; DBuf<uint32_t> d_qnib; DBuf<bg_run> d_runs;
	DBuf<uint32_t> d_best; DBuf<uint16_t> d_best16;
	DBuf<unsigned long long> d_acx_off; DBuf<uint8_t> d_post; DBuf<uint32_t> d_bad, d_cg_cnt, d_cg_first, d_cg_cache; DBuf<bg_xhit> d_xhits;   // accelerator on the device, candidate-generation scratch
	int acx_n = 0, acx_big = 0; uint32_t acx_nbad = 0, acx_clumps = 0, cg_blocks = 0; uint32_t runs_cap = 0;
	DBuf<uint16_t> d_rlen, d_rbud; DBuf<uint32_t> d_strand, d_candoff, d_cand; DBuf<unsigned long long> d_rl64, d_sl64, d_roff;   // compact strand batches
	DBuf<uint32_t> d_cls; DBuf<uint4> d_xs;                       // band-class bins of the survivors, expanded records (k_bin_*)
	DBuf<Surv> d_surv; DBuf<Res> d_res; DBuf<bg_hit> d_hits, d_hits_sorted; DBuf<uint32_t> d_scratch; uint32_t scratch_w = 1024, len_hint = 0; bool wide_possible = false;   // cells per thread of the generic band launch
	DBuf<unsigned long long> d_keys, d_keys2; DBuf<uint32_t> d_order, d_order2; DBuf<uint8_t> d_sort_tmp;
	DBuf<uint32_t> d_counters; DBuf<unsigned long long> d_cells;   // cells: [0] band, [1..4] work stats
	uint32_t *h_pinned = nullptr;                                 // 16 x u32 pinned scratch for small readbacks
	int kind = WORK_NONE;
	uint32_t nq = 0, nslots = 0, ntiles = 0; uint64_t nruns = 0, ntasks = 0;
	std::vector<uint32_t> task0;                                  // WORK_TASKS: first task index of each run
	SeedLayout SL = {0, 0, 0, 0, 0, 0, 0}; uint32_t nseed = 0;    // queries taken by k_seed
	uint32_t surv_cap = 0; bool surv_cap_forced = false;
	int last_mode = 0; std::vector<uint16_t> last_best_in; bool have_best_in = false;
	bg_stats stats;
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
	Slice sl[NSLICEBUF]; cudaStream_t copy_stream = nullptr; DBuf<uint32_t> d_first; int pipe_slices = 4, pipe_min_runs = 4096, pipe_ratio = 0;   // pipelined one-call path
	uint32_t h_counters[4] = {0, 0, 0, 0};
	bool ran = false, sorted = false;
};

static Work work_of(const bg_ctx *c) {
	Work W; W.runs = c->kind == WORK_ALL ? nullptr : c->d_runs.p; W.nruns = c->nruns; W.nq = c->nq; W.ntiles = c->ntiles;
	W.first_clump = c->first_clump; W.num_clumps = c->num_clumps; W.q_base = 0; W.run_base = 0;
	return W;
}

extern "C" void bg_default_scoring(int z, uint8_t S[256]) {
	// IUPAC code -> base set (A=1, C=2, G=4, T=8) in the reference's alphabet order
	// ". A C G T N K M R Y S W B V H D" (burst.c:166); rule of the table at burst.c:172-190 / 1310-1328
	static const uint8_t set[16] = {0, 1, 2, 4, 8, 15, 12, 3, 5, 10, 6, 9, 14, 7, 11, 13};
	for (int q = 0; q < 16; ++q) for (int r = 0; r < 16; ++r) {
		uint8_t v;
		if (!q || !r) v = 255;
		else if (q == 5 || r == 5) v = (uint8_t)(z ? 1 : 0);
		else { uint8_t i = set[q] & set[r]; v = (i == set[q] || i == set[r]) ? 0 : 1; }
		S[q * 16 + r] = v;
	}
}

extern "C" int bg_init(int device, bg_ctx **out) {
	if (!out) return fail(BG_EINVAL, "bg_init: null out");
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) { (void)cudaGetLastError(); return fail(BG_ECUDA, "bg_init: no CUDA device (%s); this engine has no CPU fallback", cudaGetErrorString(e)); }
	if (device < 0 || device >= n) return fail(BG_EINVAL, "bg_init: device %d out of range (%d devices)", device, n);
	CU(cudaSetDevice(device));
	bg_ctx *c = new bg_ctx();
	c->device = device;
	cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, device));
	c->sms = prop.multiProcessorCount;
	CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true;
	for (int i = 0; i < 4; ++i) CU(cudaEventCreate(&c->ev[i]));
	CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
	for (int i = 0; i < NSLICEBUF; ++i) { CU(cudaEventCreateWithFlags(&c->sl[i].copied, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&c->sl[i].computed, cudaEventDisableTiming)); }
	CU(cudaMallocHost((void **)&c->h_pinned, 256));
	memset(&c->stats, 0, sizeof(c->stats));
	// tuning knobs may also come from the environment (the drop-in binary has no flags for them)
	if (const char *e = getenv("BURST_B200_SEED_STAGE")) c->seed_stage = atoi(e) != 0;
	if (const char *e = getenv("BURST_B200_SEED_FILTER")) c->seed_filter = atoi(e) != 0;
	if (const char *e = getenv("BURST_B200_SEED_CHUNK")) { const int v = atoi(e); if (v >= 1 && v <= 4096) c->seed_chunk = v; }
	if (const char *e = getenv("BURST_B200_PIPE_SLICES")) { const int v = atoi(e); if (v >= 0 && v <= 64) c->pipe_slices = v; }
	if (const char *e = getenv("BURST_B200_SEED_IMPL")) c->seed_impl = atoi(e) != 0;
	if (const char *e = getenv("BURST_B200_SEED_NCH")) { const int v = atoi(e); if (v == 0 || (v >= 4 && v <= 8)) c->seed_nch = v; }
	if (const char *e = getenv("BURST_B200_SEED_VMODE")) { const int v = atoi(e); if (v == 0 || v == 1) c->seed_vmode = v; }
	if (const char *e = getenv("BURST_B200_SEED_FB")) { const int v = atoi(e); if (v == 1 || v == 2) c->seed_fb = v; }
	if (const char *e = getenv("BURST_B200_SEED_LBITS")) { const int v = atoi(e); if (v == 0 || (v >= 10 && v <= 20)) c->seed_lbits = v; }
	bg_default_scoring(1, c->S);
	*out = c;
	return bg_set_scoring(c, c->S);
}

extern "C" void bg_free(bg_ctx *c) {
	if (!c) return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	c->d_sterm.release(); c->d_db.release(); c->d_clump_off.release(); c->d_clump_len.release(); c->d_meta.release();
	c->d_packed.release(); c->d_codes.release(); c->d_qoff.release(); c->d_budget.release(); c->d_slot.release();
	c->d_qi.release(); c->d_qnib.release(); c->d_runs.release();
	c->d_best.release(); c->d_best16.release(); c->d_surv.release(); c->d_res.release(); c->d_acx_off.release(); c->d_post.release(); c->d_bad.release(); c->d_cg_cnt.release(); c->d_cg_first.release(); c->d_cg_cache.release(); c->d_xhits.release();
	c->d_cls.release(); c->d_xs.release(); c->d_rlen.release(); c->d_rbud.release(); c->d_strand.release(); c->d_candoff.release(); c->d_cand.release(); c->d_rl64.release(); c->d_sl64.release(); c->d_roff.release();
	c->d_hits.release(); c->d_hits_sorted.release(); c->d_scratch.release(); c->d_counters.release(); c->d_cells.release();
	c->d_keys.release(); c->d_keys2.release(); c->d_order.release(); c->d_order2.release(); c->d_sort_tmp.release();
	for (int i = 0; i < 4; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
	if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
	for (int i = 0; i < NSLICEBUF; ++i) { c->sl[i].release(); if (c->sl[i].copied) cudaEventDestroy(c->sl[i].copied); if (c->sl[i].computed) cudaEventDestroy(c->sl[i].computed); }
	c->d_first.release();
	if (c->h_pinned) cudaFreeHost(c->h_pinned);
	if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

extern "C" void *bg_host_alloc(uint64_t bytes) {
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); fail(BG_ENOMEM, "cudaHostAlloc(%llu bytes) failed", (unsigned long long)bytes); return nullptr; }
	return p;
}
extern "C" void bg_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" int bg_set_stream(bg_ctx *c, void *s) {
	if (!c) return fail(BG_EINVAL, "null ctx");
	if (c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
	c->stream = (cudaStream_t)s; c->own_stream = false;
	return BG_OK;
}

extern "C" int bg_set_param(bg_ctx *c, int what, int value) {
	if (!c) return fail(BG_EINVAL, "null ctx");
	if (what == BG_PARAM_SEED_FILTER) { c->seed_filter = value != 0; return BG_OK; }
	if (what == BG_PARAM_SEED_CHUNK) { if (value < 1 || value > 4096) return fail(BG_EINVAL, "bg_set_param: seed chunk %d out of range", value); c->seed_chunk = value; return BG_OK; }
	if (what == BG_PARAM_SEED_WORDS) {
		if (value && (value < 128 || value > 8192 || (value & (value - 1)))) return fail(BG_EINVAL, "bg_set_param: seed words %d must be 0 or a power of two in 128..8192", value);
		c->seed_words = value; return BG_OK;
	}
	if (what == BG_PARAM_SEED_STAGE) { c->seed_stage = value != 0; return BG_OK; }
	if (what == BG_PARAM_SEED_GROUPS) { if (value != 0 && value != 2 && value != 4 && value != 8) return fail(BG_EINVAL, "bg_set_param: seed groups %d must be 0, 2, 4 or 8", value); c->seed_groups = value; return BG_OK; }
	if (what == BG_PARAM_SEED_IMPL) { c->seed_impl = value != 0; return BG_OK; }
	if (what == BG_PARAM_SEED_NCH) { if (value != 0 && (value < 4 || value > 8)) return fail(BG_EINVAL, "bg_set_param: seed buffer chunks %d must be 0 (from the database) or 4..8", value); c->seed_nch = value; return BG_OK; }
	if (what == BG_PARAM_SEED_HSLOTS) { if (value && (value < 64 || value > 4096 || (value & (value - 1)))) return fail(BG_EINVAL, "bg_set_param: %d window-table buckets (must be 0 or a power of two 64..4096)", value); c->seed_hslots = value; return BG_OK; }
	if (what == BG_PARAM_SEED_FB) { if (value != 1 && value != 2) return fail(BG_EINVAL, "bg_set_param: filter bits per window %d must be 1 or 2", value); c->seed_fb = value; return BG_OK; }
	if (what == BG_PARAM_SEED_VMODE) { if (value != 0 && value != 1) return fail(BG_EINVAL, "bg_set_param: verification mode %d must be 0 or 1", value); c->seed_vmode = value; return BG_OK; }
	if (what == BG_PARAM_SEED_LBITS) { if (value && (value < 10 || value > 20)) return fail(BG_EINVAL, "bg_set_param: seed bitmap 2^%d bits out of range (0 = auto, 10..20)", value); c->seed_lbits = value; return BG_OK; }
	if (what == BG_PARAM_PIPE_RATIO) { if (value && (value < 10 || value > 300)) return fail(BG_EINVAL, "bg_set_param: slice ratio %d must be 0 (auto) or 10..300", value); c->pipe_ratio = value; return BG_OK; }
	if (what == BG_PARAM_PIPE_MIN_RUNS) { if (value < 1) return fail(BG_EINVAL, "bg_set_param: minimum runs per slice %d", value); c->pipe_min_runs = value; return BG_OK; }
	if (what == BG_PARAM_PIPE_SLICES) { if (value < 0 || value > 64) return fail(BG_EINVAL, "bg_set_param: pipeline slices %d out of range 0..64", value); c->pipe_slices = value; return BG_OK; }
	return fail(BG_EINVAL, "bg_set_param: unknown parameter %d", what);
}

extern "C" int bg_set_scoring(bg_ctx *c, const uint8_t S[256]) {
	if (!c || !S) return fail(BG_EINVAL, "bg_set_scoring: null argument");
	CU(cudaSetDevice(c->device));
	memcpy(c->S, S, 256);
	uint32_t st[256];
	for (int i = 0; i < 256; ++i) st[i] = (uint32_t)S[i] << 22;
	// What the seed filter needs from the table: plain bases match exactly themselves among plain bases;
	// which other reference codes can match a plain base for free decides what the scan must always verify.
	c->seed_ok = true; c->amb_add = 0x22222222u;
	for (int q = 1; q <= 4; ++q) for (int r = 0; r < 16; ++r) {
		const bool m = S[q * 16 + r] == 0;
		if (r >= 1 && r <= 4) { if (m != (q == r)) c->seed_ok = false; }
		else if (m && r == 0) c->seed_ok = false;
		else if (m && r == 5) c->amb_add = 0x33333333u;
	}
	memset(c->m16, 0, sizeof(c->m16));
	for (int q = 0; q < 16; ++q) for (int r = 0; r < 16; ++r) if (S[q * 16 + r] == 0) c->m16[q >> 1] |= 1u << (16 * (q & 1) + r);
	if (c->d_sterm.need(256)) return BG_ENOMEM;
	CU(cudaMemcpyAsync(c->d_sterm.p, st, sizeof(st), cudaMemcpyHostToDevice, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	return BG_OK;
}

extern "C" int bg_load_db(bg_ctx *c, const uint8_t *packed, const uint32_t *clump_len, uint32_t num_clumps, uint32_t first_clump) {
	if (!c || !packed || !clump_len || !num_clumps) return fail(BG_EINVAL, "bg_load_db: null/empty argument");
	CU(cudaSetDevice(c->device));
	std::vector<uint64_t> in_off(num_clumps + 1), out_off(num_clumps + 1);
	in_off[0] = out_off[0] = 0;
	for (uint32_t i = 0; i < num_clumps; ++i) {
		if (!clump_len[i]) return fail(BG_EINVAL, "bg_load_db: clump %u has length 0", i);
		in_off[i + 1] = in_off[i] + (uint64_t)((clump_len[i] + 1) / 2) * 16;
		out_off[i + 1] = out_off[i] + (uint64_t)((clump_len[i] + 31) / 32) * 16;
	}
	c->num_clumps = num_clumps; c->first_clump = first_clump;
	c->kind = WORK_NONE;
	if (c->d_db.need(out_off[num_clumps])) return BG_ENOMEM;
	if (c->d_clump_off.need(num_clumps + 1) || c->d_clump_len.need(num_clumps) || c->d_meta.need(num_clumps)) return BG_ENOMEM;
	{
		std::vector<ClumpMeta> meta(num_clumps);
		uint64_t maxb = 0;
		for (uint32_t i = 0; i < num_clumps; ++i) { meta[i].off = out_off[i]; meta[i].len = clump_len[i]; meta[i].flags = 0; maxb = std::max<uint64_t>(maxb, (out_off[i + 1] - out_off[i]) * 16); }
		c->stage_bytes = (uint32_t)std::min<uint64_t>(maxb, 8192);
		// k_seedw walks a clump in items of NCH chunks of 32 columns and probes every word of an item, valid or not: take the NCH in
		// 4..8 that wastes the fewest probes over this database (an item also costs about twelve chunks' worth of bookkeeping -- measured: 214-column clumps in two items of 4 take 1.24 ms, in one of 7 0.91 ms; ties go
		// to the larger item).  The usual sheared database has ONE clump length -- 214 columns for 100-base reads = 7 chunks.
		{
			uint64_t cost[9] = {0};
			for (uint32_t i = 0; i < num_clumps; ++i) { const uint64_t ch = (clump_len[i] + 31) / 32; for (int n = 4; n <= 8; ++n) cost[n] += (ch + n - 1) / n * (uint64_t)(n + 12); }
			int bestn = 8; for (int n = 7; n >= 4; --n) if (cost[n] < cost[bestn]) bestn = n;
			c->seed_nch_auto = bestn;
		}
		CU(cudaMemcpyAsync(c->d_meta.p, meta.data(), (size_t)num_clumps * sizeof(ClumpMeta), cudaMemcpyHostToDevice, c->stream));
		CU(cudaStreamSynchronize(c->stream));
	}
	DBuf<uint64_t> d_in_off;
	if (d_in_off.need(num_clumps + 1)) return BG_ENOMEM;
	CU(cudaMemcpyAsync(c->d_clump_off.p, out_off.data(), (num_clumps + 1) * 8, cudaMemcpyHostToDevice, c->stream));
	CU(cudaMemcpyAsync(d_in_off.p, in_off.data(), (num_clumps + 1) * 8, cudaMemcpyHostToDevice, c->stream));
	CU(cudaMemcpyAsync(c->d_clump_len.p, clump_len, num_clumps * 4, cudaMemcpyHostToDevice, c->stream));
	// stream the file-order bytes through a bounded staging buffer, transposing on the device
	const uint64_t SLAB = 256ull << 20;
	DBuf<uint8_t> stage;
	uint64_t biggest = 0;
	for (uint32_t i = 0; i < num_clumps; ++i) biggest = std::max(biggest, in_off[i + 1] - in_off[i]);
	if (stage.need(std::min<uint64_t>(std::max(SLAB, biggest), in_off[num_clumps]))) { d_in_off.release(); return BG_ENOMEM; }
	uint32_t i = 0;
	while (i < num_clumps) {
		uint32_t j = i;
		while (j < num_clumps && in_off[j + 1] - in_off[i] <= stage.cap && j - i < (1u << 30)) ++j;
		if (j == i) { stage.release(); d_in_off.release(); return fail(BG_EINVAL, "bg_load_db: clump %u larger than staging", i); }
		uint64_t bytes = in_off[j] - in_off[i];
		CU(cudaMemcpyAsync(stage.p, packed + in_off[i], bytes, cudaMemcpyHostToDevice, c->stream));
		k_relayout<<<j - i, 128, 0, c->stream>>>(stage.p, d_in_off.p, c->d_clump_off.p, c->d_clump_len.p, c->d_db.p, i, in_off[i], (uint32_t *)c->d_meta.p);
		CU(cudaGetLastError());
		CU(cudaStreamSynchronize(c->stream));
		i = j;
	}
	stage.release(); d_in_off.release();
	return BG_OK;
}

// per-clump counters of the resident candidate-generation blocks (the accelerator names clumps of the whole database); private to a context
static int acx_scratch(bg_ctx *c) {
	c->acx_clumps = c->first_clump + c->num_clumps;
	uint32_t blocks = (uint32_t)c->sms * 4;
	while (blocks > (uint32_t)c->sms && (size_t)blocks * c->acx_clumps * 12 > (4ull << 30)) blocks -= (uint32_t)c->sms;
	c->cg_blocks = blocks;
	const size_t ne = (size_t)blocks * c->acx_clumps;
	if (c->d_cg_cnt.need(ne) || c->d_cg_first.need(ne) || c->d_cg_cache.need(ne)) return BG_ENOMEM;
	CU(cudaMemsetAsync(c->d_cg_cnt.p, 0, ne * 4, c->stream));
	CU(cudaMemsetAsync(c->d_cg_first.p, 0xFF, ne * 4, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	return BG_OK;
}

// A second context on the same device that uses the database (and accelerator) already loaded by `src` instead of a copy of its own:
// two contexts = two batches in flight on one GPU (the copies of one behind the kernels of the other), one host thread each.
// `src` must outlive `ctx` or be freed after it; loading a database of its own into `ctx` later simply drops the borrowed one.
extern "C" int bg_share_db(bg_ctx *c, bg_ctx *src) {
	if (!c || !src || c == src) return fail(BG_EINVAL, "bg_share_db: null or identical contexts");
	if (c->device != src->device) return fail(BG_EINVAL, "bg_share_db: contexts on different devices (%d, %d)", c->device, src->device);
	if (!src->num_clumps) return fail(BG_EINVAL, "bg_share_db: the source context holds no database");
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(src->stream));                        // the source's load is complete before anyone reads it from another stream
	c->d_db.borrow(src->d_db); c->d_clump_off.borrow(src->d_clump_off); c->d_clump_len.borrow(src->d_clump_len); c->d_meta.borrow(src->d_meta);
	c->num_clumps = src->num_clumps; c->first_clump = src->first_clump; c->stage_bytes = src->stage_bytes; c->seed_nch_auto = src->seed_nch_auto;
	c->d_acx_off.borrow(src->d_acx_off); c->d_post.borrow(src->d_post); c->d_bad.borrow(src->d_bad);
	c->acx_n = 0; c->acx_big = src->acx_big; c->acx_nbad = src->acx_nbad;
	if (src->acx_n) { int rc = acx_scratch(c); if (rc) return rc; c->acx_n = src->acx_n; }
	c->kind = WORK_NONE;
	if (!src->gate) {
		src->gate = std::make_shared<Gate>();
		cudaEvent_t e = nullptr; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); src->gate->ev.push_back(e); src->gate_done = e;
	}
	if (c->gate != src->gate) {
		std::lock_guard<std::mutex> g(src->gate->m);
		cudaEvent_t e = nullptr; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); src->gate->ev.push_back(e);
		c->gate = src->gate; c->gate_done = e;
	}
	return BG_OK;
}

// Window layout for the batch from the histogram of stretch lengths (plen = len / (k+1), capped at 31):
// the longest window (<= 16 bases) that ~all seedable queries can afford, probing every 8 columns when
// that still leaves >= 14 bases, else every 4.
static SeedLayout choose_layout(const bg_ctx *c, const uint32_t hist[32], uint32_t nq) {
	SeedLayout L = {0, 0, 0, 0, 0, 0, 0};
	if (!c->seed_filter || !c->seed_ok) return L;
	uint64_t total = 0;
	for (uint32_t p = 11; p < 32; ++p) total += hist[p];
	if (!total || total * 4 < nq) return L;                      // too few queries could use it
	uint64_t acc = 0; uint32_t P = 11;
	for (uint32_t p = 31; p >= 11; --p) { acc += hist[p]; if (acc * 20 >= total * 19) { P = p; break; } }
	const uint32_t w8 = P >= 15 ? std::min<uint32_t>(16, P - 7) : 0, w4 = std::min<uint32_t>(16, P - 3);
	if (w8 >= 14) { L.stride = 8; L.w = w8; } else { L.stride = 4; L.w = w4; }
	L.hm = L.w > 8 ? 0xFFFFFFFFu << (4 * (16 - L.w)) : 0u;
	L.amb_add = c->amb_add; L.np_max = 128 / L.stride;
	return L;
}

// Query bases [b0, b1) of the caller's code array -> `codes` on the device (one byte per base), through the
// nibble-packed staging buffer when the caller's array is BG_Q_PACKED4.
static int copy_codes(const bg_queries *Q, uint64_t b0, uint64_t b1, DBuf<uint8_t> &packed, DBuf<uint8_t> &codes, cudaStream_t copy, cudaStream_t unpack, cudaEvent_t copied) {
	const uint64_t n = b1 - b0;
	if (codes.need(n + 32)) return BG_ENOMEM;
	if (!(Q->flags & BG_Q_PACKED4)) { CU(cudaMemcpyAsync(codes.p, Q->codes + b0, n, cudaMemcpyHostToDevice, copy)); return BG_OK; }
	const uint64_t y0 = b0 >> 1, y1 = (b1 + 1) >> 1;
	if (packed.need(y1 - y0 + 32)) return BG_ENOMEM;
	CU(cudaMemcpyAsync(packed.p, Q->codes + y0, y1 - y0, cudaMemcpyHostToDevice, copy));
	if (copy != unpack) { CU(cudaEventRecord(copied, copy)); CU(cudaStreamWaitEvent(unpack, copied, 0)); }
	if (n) k_unpack4<<<(unsigned)((n + 16 * 256 - 1) / (16 * 256)), 256, 0, unpack>>>(packed.p, (uint32_t)(b0 & 1), n, codes.p);
	CU(cudaGetLastError());
	return BG_OK;
}

// queries -> device, QInfo + tables
// query length the k_extend staging slots are sized for, from the longest of a sample of the batch (a longer query still works: it reads global memory)
static bool long_queries(uint64_t sampled_max) { return sampled_max > 400; }   // budgets above ~8 at -i 0.98: generic bands become likely
static uint32_t stage_len(uint64_t sampled_max) {
	static const int pad = getenv("BURST_B200_STAGE_PAD") ? atoi(getenv("BURST_B200_STAGE_PAD")) : 0;
	static const uint64_t cap = getenv("BURST_B200_STAGE_CAP") ? (uint64_t)atoi(getenv("BURST_B200_STAGE_CAP")) : 320;
	return (uint32_t)std::min<uint64_t>(cap, ((sampled_max + 7) & ~7ull) + (uint64_t)pad);   // (320: the slots of 128 threads still fit; a batch that also holds longer queries keeps staging its short ones)
}

static int upload_queries(bg_ctx *c, const bg_queries *Q) {
	if (!c->num_clumps) return fail(BG_EINVAL, "bg_batch_upload: no database loaded");
	if (!Q->nq) return fail(BG_EINVAL, "bg_batch_upload: empty query batch");
	if (!Q->codes || !Q->offset || !Q->budget || !Q->slot) return fail(BG_EINVAL, "bg_batch_upload: null query array");
	const uint32_t nq = Q->nq;
	const uint64_t ncodes = Q->offset[nq];
	if (Q->flags & ~(uint32_t)BG_Q_PACKED4) return fail(BG_EINVAL, "bg_batch_upload: unknown query flags %u", Q->flags);
	if (c->d_qoff.need(nq + 1) || c->d_budget.need(nq) || c->d_slot.need(nq) || c->d_qi.need(nq) ||
	    c->d_qnib.need(ncodes / 8 + 3ull * nq + 8) || c->d_best.need(Q->nslots) || c->d_best16.need(Q->nslots) ||
	    c->d_counters.need(64) || c->d_cells.need(8)) return BG_ENOMEM;
	{ int rc = copy_codes(Q, Q->offset[0], ncodes, c->d_packed, c->d_codes, c->stream, c->stream, c->ev[3]); if (rc) return rc; }
	CU(cudaMemcpyAsync(c->d_qoff.p, Q->offset, (size_t)(nq + 1) * 8, cudaMemcpyHostToDevice, c->stream));
	CU(cudaMemcpyAsync(c->d_budget.p, Q->budget, (size_t)nq * 2, cudaMemcpyHostToDevice, c->stream));
	CU(cudaMemcpyAsync(c->d_slot.p, Q->slot, (size_t)nq * 4, cudaMemcpyHostToDevice, c->stream));
	CU(cudaMemsetAsync(c->d_counters.p, 0, 256, c->stream));          // [0..3] counters, [9] seed queries, [16..47] stretch-length histogram
	k_qinfo<<<(nq + 255) / 256, 256, 0, c->stream>>>(c->d_qoff.p, c->d_budget.p, c->d_slot.p, nq, Q->nslots, c->d_qi.p, c->d_counters.p + 16, c->d_counters.p, Q->offset[0]);
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(c->h_pinned, c->d_counters.p, 256, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	if (c->h_pinned[C_ERR]) {
		uint32_t q = c->h_pinned[C_ERR] - 1;
		return fail(BG_EINVAL, "bg_batch_upload: query %u is malformed (length %llu, budget %u (max 254, burst.c:3076), slot %u of %u)", q,
			(unsigned long long)(Q->offset[q + 1] - Q->offset[q]), Q->budget[q], Q->slot[q], Q->nslots);
	}
	c->SL = choose_layout(c, c->h_pinned + 16, nq);
	{ uint64_t mx = 0; const uint32_t stp = std::max<uint32_t>(1, nq / 8192); for (uint32_t q = 0; q < nq; q += stp) mx = std::max<uint64_t>(mx, Q->offset[q + 1] - Q->offset[q]); c->mstage = stage_len(mx); c->wide_possible = long_queries(mx); c->len_hint = (uint32_t)mx; }
	k_qprep<<<(nq + 127) / 128, 128, 0, c->stream>>>(c->d_codes.p, c->d_qi.p, nq, c->SL, c->d_qnib.p, c->d_counters.p + 9, nullptr);
	CU(cudaGetLastError());
	c->nq = nq; c->nslots = Q->nslots;
	return BG_OK;
}

static void seed_sizes(bg_ctx *c, SeedLayout &SL, uint32_t stretches_mean, uint32_t stretches_max, uint32_t &npmax);

static int finish_upload(bg_ctx *c) {
	CU(cudaMemcpyAsync(c->h_pinned, c->d_counters.p, 64, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));     // the caller's host buffers may be reused after this returns
	if (c->h_pinned[C_ERR]) { c->kind = WORK_NONE; return fail(BG_EINVAL, "bg_batch_upload_runs: run %u is malformed (nq must be 1..%d and query0+nq within the batch)", c->h_pinned[C_ERR] & 0x7FFFFFFF, BG_RUN_MAX); }
	c->nseed = c->h_pinned[9];
	if (c->nseed) seed_sizes(c, c->SL, (c->h_pinned[10] + c->nseed - 1) / c->nseed, c->h_pinned[11], c->seed_npmax);
	memset(&c->stats, 0, sizeof(c->stats));
	if (!c->surv_cap) c->surv_cap = 1u << 20;
	uint64_t want = std::min<uint64_t>(c->ntasks * 16, std::max<uint64_t>(c->surv_cap, std::min<uint64_t>(4ull * c->nq + c->ntasks / 8, 1ull << 26)));   // (at most 64 M entries up front: 6 GB of lists; more only when a batch proves to need it)
	want = std::max<uint64_t>(want, 1024);
	if (c->surv_cap_forced) { want = c->surv_cap; c->surv_cap_forced = false; }      // (bg_set_surv_cap: the next batch starts from exactly that size)
	if (want > c->surv_cap || !c->d_surv.p) c->surv_cap = (uint32_t)std::min<uint64_t>(want, 0xFFFFFFF0ull);
	if (c->d_surv.need(c->surv_cap) || c->d_xs.need((size_t)c->surv_cap * 3) || c->d_cls.need(64) || c->d_res.need(c->surv_cap) || c->d_hits.need(c->surv_cap) || c->d_keys.need(c->surv_cap)) return BG_ENOMEM;
	c->ran = false; c->sorted = false;
	return BG_OK;
}

extern "C" int bg_batch_upload_runs(bg_ctx *c, const bg_queries *Q, const bg_run *runs, uint64_t nruns) {
	if (!c || !Q || !runs) return fail(BG_EINVAL, "bg_batch_upload_runs: null argument");
	if (nruns >= (1ull << 28)) return fail(BG_EINVAL, "bg_batch_upload_runs: %llu runs in one batch (limit 2^28-1); split the query batch", (unsigned long long)nruns);
	CU(cudaSetDevice(c->device));
	c->kind = WORK_NONE;
	int rc = upload_queries(c, Q); if (rc) return rc;
	if (c->d_runs.need(nruns + 1)) return BG_ENOMEM;
	CU(cudaMemcpyAsync(c->d_runs.p, runs, nruns * sizeof(bg_run), cudaMemcpyHostToDevice, c->stream));
	if (nruns) k_check_runs<<<(unsigned)((nruns + 255) / 256), 256, 0, c->stream>>>(c->d_runs.p, nruns, 0, c->nq, c->d_counters.p);
	CU(cudaGetLastError());
	c->kind = WORK_RUNS; c->nruns = nruns; c->ntasks = nruns * BG_RUN_MAX; c->ntiles = 0;
	return finish_upload(c);
}

extern "C" int bg_batch_upload(bg_ctx *c, const bg_queries *Q, const bg_task *tasks, uint64_t ntasks) {
	if (!c || !Q) return fail(BG_EINVAL, "bg_batch_upload: null argument");
	CU(cudaSetDevice(c->device));
	c->kind = WORK_NONE;
	if (!tasks) {                                                 // all-vs-all, reference order clump-major (burst.c:4344, 4365)
		int rc = upload_queries(c, Q); if (rc) return rc;
		c->ntiles = (Q->nq + BG_RUN_MAX - 1) / BG_RUN_MAX;
		c->nruns = (uint64_t)c->ntiles * c->num_clumps;
		c->ntasks = (uint64_t)Q->nq * c->num_clumps;
		if (c->ntasks >= (1ull << 32) || c->nruns >= (1ull << 28))
			return fail(BG_EINVAL, "bg_batch_upload: %llu tasks in one all-vs-all batch (limits 2^32-1 tasks, 2^28-1 clump x 16-query tiles); split the query batch", (unsigned long long)c->ntasks);
		c->kind = WORK_ALL;
		return finish_upload(c);
	}
	if (ntasks >= (1ull << 32)) return fail(BG_EINVAL, "bg_batch_upload: %llu tasks in one batch (limit 2^32-1); split the query batch", (unsigned long long)ntasks);
	// coalesce the task list into runs: consecutive tasks on one clump with consecutive query ids
	std::vector<bg_run> runs; runs.reserve(ntasks / 8 + 16);
	c->task0.clear(); c->task0.reserve(ntasks / 8 + 16);
	for (uint64_t t = 0; t < ntasks; ++t) {
		if (tasks[t].query >= Q->nq) return fail(BG_EINVAL, "bg_batch_upload: task %llu names query %u of %u", (unsigned long long)t, tasks[t].query, Q->nq);
		if (!runs.empty()) {
			bg_run &R = runs.back();
			if (R.clump == tasks[t].clump && R.nq < BG_RUN_MAX && tasks[t].query == R.query0 + R.nq) { ++R.nq; continue; }
		}
		runs.push_back(bg_run{tasks[t].clump, tasks[t].query, 1}); c->task0.push_back((uint32_t)t);
	}
	if (runs.size() >= (1ull << 28)) return fail(BG_EINVAL, "bg_batch_upload: task list needs %zu runs (limit 2^28-1)", runs.size());
	int rc = upload_queries(c, Q); if (rc) return rc;
	if (c->d_runs.need(runs.size() + 1)) return BG_ENOMEM;
	CU(cudaMemcpyAsync(c->d_runs.p, runs.data(), runs.size() * sizeof(bg_run), cudaMemcpyHostToDevice, c->stream));
	c->kind = WORK_TASKS; c->nruns = runs.size(); c->ntasks = ntasks; c->ntiles = 0;
	return finish_upload(c);                                      // synchronises: `runs` may go out of scope
}

// Device view of one uploaded batch (or one slice of a pipelined one).
struct BatchDev { const uint8_t *codes; const uint32_t *qnib; const QInfo *qi; Work W; };

static void seed_sizes(bg_ctx *c, SeedLayout &SL, uint32_t stretches_mean, uint32_t stretches_max, uint32_t &npmax) {
	// Bloom filter size: 1-2 words per window of a full bunch (16 queries x mean stretches x stride): a false positive
	// costs the owning thread one hash-table probe, a larger filter costs occupancy (measured: profiles/)
	const uint64_t windows = (uint64_t)BG_RUN_MAX * SL.stride * stretches_mean;
	uint32_t words = 256;
	while (words < 4096 && words < windows) words <<= 1;
	if (c->seed_words) words = (uint32_t)c->seed_words;
	SL.words = words; SL.shw = 32; for (uint32_t w = words; w > 1; w >>= 1) --SL.shw;
	npmax = std::min<uint32_t>(std::max<uint32_t>(stretches_max, 1), 128 / SL.stride);   // window table rows per query
}

// the warp form of the seed filter: table sizes from the batch, grid = what is resident at once
static int launch_seedw(bg_ctx *c, cudaStream_t st, const BatchDev &B, const SeedLayout &SL, uint32_t npmax) {
	SeedWArgs S;
	const int nch = c->seed_nch ? c->seed_nch : c->seed_nch_auto;
	S.db = c->d_db.p; S.meta = c->d_meta.p; S.qi = B.qi; S.qnib = B.qnib; S.W = B.W; S.SL = SL; S.nwork = B.W.nruns;
	uint32_t chunk = 1; while (chunk * 2 <= (uint32_t)c->seed_chunk * 8 && chunk < SEEDW_SPLIT) chunk <<= 1;   // a power of two that divides SEEDW_SPLIT (default 64 runs)
	S.chunk = chunk; S.npmax = npmax;
	const uint32_t ne = BG_RUN_MAX * npmax * SL.stride;                 // most windows a bunch can hold
	// buckets of the chained window table: about 2/3 of the windows a full bunch can hold (chains of 1-2; more buckets cost shared memory)
	S.hslots = 64; while (S.hslots * 3 < ne * 2) S.hslots <<= 1;
	if (c->seed_hslots) S.hslots = (uint32_t)c->seed_hslots;
	// filter: FB bits per window inside one 32-bit word.  Start from ~64 bits per window of a typical full bunch, then give up one
	// power of two when that lets one more block live on an SM (measured on the bench shape: 2 bits in 2^14 at 4 blocks/SM beats
	// 1 bit in 2^15 at 3 blocks/SM by 7 %, profiles/r2_tune_seedw.txt); a false positive costs one bucket probe
	auto smem_of = [&](uint32_t lb) { return (16 + (size_t)SEEDW_WARPS * seedw_warp_words(lb, S.hslots, npmax, SL.stride, (uint32_t)nch)) * sizeof(uint32_t); };
	auto blocks_of = [&](uint32_t lb) { return std::min<size_t>(4, (size_t)(227 * 1024) / (smem_of(lb) + 1024)); };
	uint32_t lbits = 12; while (lbits < 17 && (1u << lbits) < 64u * SL.words) ++lbits;
	if (c->seed_lbits) lbits = (uint32_t)c->seed_lbits;
	else if (lbits > 12 && (1u << (lbits - 1)) >= 32u * ne && blocks_of(lbits - 1) > blocks_of(lbits)) --lbits;
	while (lbits > 10 && smem_of(lbits) > 200 * 1024) --lbits;
	S.lbits = lbits;
	const size_t smem = (16 + (size_t)SEEDW_WARPS * seedw_warp_words(lbits, S.hslots, npmax, SL.stride, (uint32_t)nch)) * sizeof(uint32_t);
	if (smem > 220 * 1024) return fail(BG_EINVAL, "seed filter tables do not fit shared memory (stretches %u, stride %u)", npmax, SL.stride);
	S.surv = c->d_surv.p; S.surv_cap = c->surv_cap; S.counters = c->d_counters.p;
	memcpy(S.m16, c->m16, sizeof(S.m16));
	void (*kern)(SeedWArgs);
	#define SEEDW_PICK(NCH, FB, VM) (SL.stride == 8 ? (SL.w == 16 ? k_seedw<8, true, NCH, FB, VM> : k_seedw<8, false, NCH, FB, VM>) : (SL.w == 16 ? k_seedw<4, true, NCH, FB, VM> : k_seedw<4, false, NCH, FB, VM>))
	if (c->seed_fb != 2) kern = nch <= 4 ? SEEDW_PICK(4, 1, 0) : SEEDW_PICK(8, 1, 0);     // (one bit per window is a tuning setting: two item sizes only)
	else if (c->seed_vmode == 0) kern = nch == 4 ? SEEDW_PICK(4, 2, 0) : nch == 5 ? SEEDW_PICK(5, 2, 0) : nch == 6 ? SEEDW_PICK(6, 2, 0) : nch == 7 ? SEEDW_PICK(7, 2, 0) : SEEDW_PICK(8, 2, 0);
	else kern = nch == 4 ? SEEDW_PICK(4, 2, 1) : nch == 5 ? SEEDW_PICK(5, 2, 1) : nch == 6 ? SEEDW_PICK(6, 2, 1) : nch == 7 ? SEEDW_PICK(7, 2, 1) : SEEDW_PICK(8, 2, 1);
	#undef SEEDW_PICK
	if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int bps = 0;
	CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, SEEDW_WARPS * 32, smem));
	if (bps < 1) bps = 1;
	if (getenv("BURST_B200_DEBUG")) fprintf(stderr, "[k_seedw] lbits %u hslots %u nch %d smem %zu B/block, %d blocks/SM\n", lbits, S.hslots, nch, smem, bps);
	const uint64_t nchunk = (S.nwork + chunk - 1) / chunk;
	const uint64_t blocks = std::min<uint64_t>((nchunk + SEEDW_WARPS - 1) / SEEDW_WARPS, (uint64_t)c->sms * bps);
	kern<<<(unsigned)blocks, SEEDW_WARPS * 32, smem, st>>>(S);
	CU(cudaGetLastError());
	return BG_OK;
}

static int launch_filters(bg_ctx *c, cudaStream_t st, const BatchDev &B, const SeedLayout &SL, uint32_t npmax, bool seed, bool filter, const uint32_t *todo) {
	if (B.W.nruns && seed && c->seed_impl) { int rc = launch_seedw(c, st, B, SL, npmax); if (rc) return rc; }
	else if (B.W.nruns && seed) {
		SeedArgs S;
		S.db = c->d_db.p; S.meta = c->d_meta.p; S.qi = B.qi;
		S.qnib = B.qnib; S.W = B.W; S.SL = SL; S.nwork = B.W.nruns; S.chunk = (uint32_t)c->seed_chunk;
		S.npmax = npmax; S.stage = c->seed_stage ? c->stage_bytes : 0;
		S.hb = 64; while (S.hb < 16 * npmax * SL.stride) S.hb <<= 1;
		S.surv = c->d_surv.p; S.surv_cap = c->surv_cap; S.counters = c->d_counters.p;
		memcpy(S.m16, c->m16, sizeof(S.m16));
		// groups (runs per round) per block: the count that keeps the most runs in flight per SM
		// a block stalls as a whole on its barriers and table builds, so small blocks (4 groups = 2 warps) hide latency best
		// (measured, profiles/): 4 unless that leaves fewer than two blocks per SM
		uint32_t gpb = 4;
		while (gpb > 2 && (size_t)seed_smem(SL.words, S.hb, npmax, SL.stride, S.stage, gpb).total * 4 > 110 * 1024) gpb >>= 1;
		if (c->seed_groups) gpb = (uint32_t)c->seed_groups;
		if ((size_t)seed_smem(SL.words, S.hb, npmax, SL.stride, S.stage, gpb).total * 4 > 220 * 1024)
			return fail(BG_EINVAL, "seed filter tables do not fit shared memory (words %u, stretches %u)", SL.words, npmax);
		const uint64_t per_block = (uint64_t)gpb * S.chunk, blocks = (B.W.nruns + per_block - 1) / per_block;
		const size_t smem = (size_t)seed_smem(SL.words, S.hb, npmax, SL.stride, S.stage, gpb).total * sizeof(uint32_t);
		void (*kern)(SeedArgs) = SL.stride == 8 ? (SL.w == 16 ? k_seed<8, true> : k_seed<8, false>) : (SL.w == 16 ? k_seed<4, true> : k_seed<4, false>);
		if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		kern<<<(unsigned)blocks, gpb * 16, smem, st>>>(S);
		CU(cudaGetLastError());
	}
	if (B.W.nruns && filter) {
		FilterArgs F;
		F.db = c->d_db.p; F.clump_off = c->d_clump_off.p; F.clump_len = c->d_clump_len.p; F.qi = B.qi;
		F.codes = B.codes; F.Sterm = c->d_sterm.p; F.W = B.W; F.surv = c->d_surv.p; F.surv_cap = c->surv_cap; F.counters = c->d_counters.p; F.c16 = 16;
		F.todo = todo;
		const uint64_t blocks = std::min<uint64_t>(B.W.nruns * 2, (uint64_t)c->sms * 16);   // 8 task ids per group, groups strided over a resident grid
		// one instance per prefix length; each skips the queries of the others (and all return at once when k_seed took every query)
		k_filter<1><<<(unsigned)blocks, 128, 0, st>>>(F); k_filter<2><<<(unsigned)blocks, 128, 0, st>>>(F); k_filter<4><<<(unsigned)blocks, 128, 0, st>>>(F);
		k_filter<8><<<(unsigned)blocks, 128, 0, st>>>(F); k_filter<16><<<(unsigned)blocks, 128, 0, st>>>(F); k_filter<32><<<(unsigned)blocks, 128, 0, st>>>(F);
		CU(cudaGetLastError());
	}
	return BG_OK;
}

static int launch_extend(bg_ctx *c, cudaStream_t st, const BatchDev &B, int mode, const uint32_t *first) {
	// survivors of this launch ([*first, count)) binned by band class, then one sweep per class over its own range
	CU(cudaMemsetAsync(c->d_cls.p, 0, 64 * sizeof(uint32_t), st));
	k_bin_count<<<(unsigned)c->sms * 4, 256, 0, st>>>(c->d_surv.p, c->d_counters.p, c->surv_cap, first, c->d_cls.p);
	k_bin_offsets<<<1, 32, 0, st>>>(c->d_counters.p, c->surv_cap, first, c->d_cls.p);
	k_bin_scatter<<<(unsigned)c->sms * 8, 256, 0, st>>>(c->d_surv.p, c->d_counters.p, c->surv_cap, first, c->d_cls.p, c->d_xs.p, B.W, B.qi, c->d_meta.p);
	ExtendArgs E;
	E.dbw = (const uint32_t *)c->d_db.p; E.meta = c->d_meta.p;
	E.codes = B.codes; E.qnib = B.qnib; E.qi = B.qi; E.W = B.W;
	E.surv = c->d_surv.p; E.surv_cap = c->surv_cap; E.counters = c->d_counters.p; E.cls = c->d_cls.p; E.xs = c->d_xs.p; E.res = c->d_res.p;
	E.best = c->d_best.p; E.Sterm = c->d_sterm.p; E.scratch = nullptr; E.scratch_w = c->scratch_w;
	E.band_cells = c->d_cells.p; E.mode = mode;
	// staging slot of a thread: the reference pieces and packed query words of one survivor of up to `mstage` bases (longer ones read
	// global memory directly); two slots per thread, 128 threads per block
	const uint32_t ms = c->mstage ? c->mstage : 128;
	static const bool no_stage = getenv("BURST_B200_EXT_STAGE") && atoi(getenv("BURST_B200_EXT_STAGE")) == 0;   // debugging: sweep straight out of global memory
	E.qp_stage = ((ms + 7) / 8 + 3 + 3) / 4;
	size_t smem[NCLASS]; unsigned grid[NCLASS];
	for (int k = 0; k < NCLASS; ++k) {
		const int wb = class_width(k);
		E.np_stage[k] = wb ? (ms + wb + 17 + 31) / 32 + 1 : 0;
		smem[k] = wb ? (size_t)128 * 2 * (E.np_stage[k] + E.qp_stage) * 16 : 0;
		if (smem[k] > 100 * 1024 || no_stage) { E.np_stage[k] = 0; smem[k] = 0; }            // too long for staging: direct reads
		grid[k] = (unsigned)c->sms * (wb && wb <= 32 ? 12 : 6);
	}
	{	// a thread prefetches its NEXT survivor while it sweeps the current one: launch what is resident at once, no more
		int occ[NCLASS] = {0};
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], k_extend<5>, 128, smem[0]); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], k_extend<8>, 128, smem[1]);
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], k_extend<12>, 128, smem[2]); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[3], k_extend<16>, 128, smem[3]);
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[4], k_extend<24>, 128, smem[4]); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[5], k_extend<32>, 128, smem[5]);
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[6], k_extend<48>, 128, smem[6]); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[7], k_extend<64>, 128, smem[7]);
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[8], k_extend<0>, 128, smem[8]);
		// resident blocks per SM: three (12 warps) measured fastest on the bench shape -- 0.495 ms against 0.52 / 0.55 ms at four / five
		// (profiles/r2_ab_extend.txt): each thread keeps more survivors in its private pipeline and the unrolled band rows of
		// twelve warps still fit the instruction cache
		static const int bps_cap = getenv("BURST_B200_EXT_BPS") ? atoi(getenv("BURST_B200_EXT_BPS")) : 3;
		for (int k = 0; k < NCLASS; ++k) if (occ[k] > 0) grid[k] = std::min<unsigned>(grid[k], (unsigned)c->sms * (unsigned)(bps_cap > 0 ? std::min(bps_cap, occ[k]) : occ[k]));
		(void)cudaGetLastError();
	}
	if (c->wide_possible) {   // long queries: the filters have counted the widest band by now -- size the scratch for it instead of finding out after a wasted sweep
		uint32_t widest = 0;
		CU(cudaMemcpyAsync(&widest, c->d_counters.p + C_SCRATCH, 4, cudaMemcpyDeviceToHost, st));
		CU(cudaStreamSynchronize(st));
		if (widest > c->scratch_w) c->scratch_w = widest + 64;
		if (widest > 416 && c->len_hint + 2 > c->scratch_w) c->scratch_w = c->len_hint + 64;   // the striped sweep of a band too wide for shared memory keeps one value per query row
	}
	{	// the generic class keeps its bands in global scratch: scratch_w cells per thread, at most 4 GB in all
		const unsigned most = (unsigned)std::max<size_t>(1, ((size_t)1 << 30) / ((size_t)128 * c->scratch_w));
		grid[8] = std::min(std::min(grid[8], most), (unsigned)c->sms);
		if (c->d_scratch.need((size_t)grid[8] * 128 * c->scratch_w)) return BG_ENOMEM;
		E.scratch = c->d_scratch.p;
		static const uint32_t smem_w_cap = getenv("BURST_B200_SMEM_W") ? (uint32_t)atoi(getenv("BURST_B200_SMEM_W")) : 416u;
		static const uint32_t gen_mode = getenv("BURST_B200_GEN_MODE") ? (uint32_t)atoi(getenv("BURST_B200_GEN_MODE")) : 1u;   // 1: reference code fetched per cell, one cell ahead (5.6 s against 8.1 s for a cached word on the manuscript data set: lanes reload at different cells and diverge)
		E.gen_mode = gen_mode;
		E.smem_w = std::min<uint32_t>(c->scratch_w, std::min<uint32_t>(smem_w_cap, 416));              // 416 cells x 128 threads x 4 B = 208 KB: one block per SM
		smem[8] = (size_t)E.smem_w * 128 * 4;
	}
	#define EXT_LAUNCH(WM, K) do { if (smem[K] > 32 * 1024) CU(cudaFuncSetAttribute(k_extend<WM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem[K]));   /* (the kernel also has 1 KB of static shared memory) */ \
		k_extend<WM><<<grid[K], 128, smem[K], st>>>(E); } while (0)
	EXT_LAUNCH(5, 0); EXT_LAUNCH(8, 1); EXT_LAUNCH(12, 2); EXT_LAUNCH(16, 3); EXT_LAUNCH(24, 4); EXT_LAUNCH(32, 5); EXT_LAUNCH(48, 6); EXT_LAUNCH(64, 7); EXT_LAUNCH(0, 8);
	#undef EXT_LAUNCH
	if (getenv("BURST_B200_DEBUG")) {                                   // class sizes, staging slots and the launch status of the sweep
		uint32_t h[64]; cudaError_t e1 = cudaStreamSynchronize(st); cudaMemcpy(h, c->d_cls.p, sizeof(h), cudaMemcpyDeviceToHost);
		fprintf(stderr, "[burst_b200] extend: mstage %u qp %u sync=%s last=%s |", ms, E.qp_stage, cudaGetErrorString(e1), cudaGetErrorString(cudaPeekAtLastError()));
		for (int k = 0; k < NCLASS; ++k) fprintf(stderr, " c%d[n=%u np=%u smem=%zu]", k, h[k], E.np_stage[k], smem[k]);
		unsigned long long cells = 0; uint32_t cn[4] = {0, 0, 0, 0};
		cudaMemcpy(&cells, c->d_cells.p, 8, cudaMemcpyDeviceToHost); cudaMemcpy(cn, c->d_counters.p, 16, cudaMemcpyDeviceToHost);
		fprintf(stderr, " | band cells so far %llu, widest generic band %u, survivors so far %u\n", cells, cn[C_SCRATCH], cn[C_SURV]);
	}
	CU(cudaGetLastError());
	return BG_OK;
}

static int run_extend(bg_ctx *c, int mode, const uint16_t *best_in) {
	CU(cudaSetDevice(c->device));
	c->last_mode = mode; c->have_best_in = best_in != nullptr;
	if (best_in) {
		if (best_in != c->last_best_in.data()) c->last_best_in.assign(best_in, best_in + c->nslots);
		CU(cudaMemcpyAsync(c->d_best16.p, c->last_best_in.data(), c->nslots * 2, cudaMemcpyHostToDevice, c->stream));
	}
	k_init_best<<<(c->nslots + 255) / 256, 256, 0, c->stream>>>(c->d_best.p, best_in ? c->d_best16.p : nullptr, c->nslots);
	CU(cudaMemsetAsync(c->d_counters.p, 0, 16, c->stream));
	CU(cudaMemsetAsync(c->d_cells.p, 0, 8, c->stream));
	CU(cudaEventRecord(c->ev[0], c->stream));
	BatchDev B; B.codes = c->d_codes.p; B.qnib = c->d_qnib.p; B.qi = c->d_qi.p; B.W = work_of(c);
	int rc = launch_filters(c, c->stream, B, c->SL, c->seed_npmax, c->nseed != 0, c->nseed < c->nq, nullptr); if (rc) return rc;
	CU(cudaEventRecord(c->ev[1], c->stream));
	rc = launch_extend(c, c->stream, B, mode, nullptr); if (rc) return rc;
	CU(cudaEventRecord(c->ev[2], c->stream));
	return BG_OK;
}

static int run_select(bg_ctx *c, int mode) {
	CU(cudaSetDevice(c->device));
	k_select<<<(unsigned)c->sms * 4, 256, 0, c->stream>>>(c->d_surv.p, c->d_res.p, c->d_best.p, c->d_counters.p,
		c->surv_cap, c->d_hits.p, c->d_keys.p, mode);
	CU(cudaGetLastError());
	CU(cudaEventRecord(c->ev[3], c->stream));
	c->ran = true; c->sorted = false;
	return BG_OK;
}

// The split form hands the per-slot minima to an external all-reduce before the selection: they must be final when this returns.
// A survivor list (or band scratch) that overflowed would leave them computed from a truncated list, so the overflow is settled
// HERE -- grow, redo filter + extend -- and not at download time as in the one-call form.
extern "C" int bg_batch_run_extend(bg_ctx *c, int mode, const uint16_t *best_in) {
	if (!c || c->kind == WORK_NONE) return fail(BG_EINVAL, "bg_batch_run: no batch uploaded");
	for (int attempt = 0; attempt < 5; ++attempt) {
		int rc = run_extend(c, mode, best_in); if (rc) return rc;
		CU(cudaMemcpyAsync(c->h_pinned, c->d_counters.p, 16, cudaMemcpyDeviceToHost, c->stream));
		CU(cudaStreamSynchronize(c->stream));
		const bool grow_s = c->h_pinned[C_SURV] > c->surv_cap, grow_g = c->h_pinned[C_SCRATCH] > c->scratch_w;
		if (!grow_s && !grow_g) return BG_OK;
		if (grow_s) {
			c->surv_cap = c->h_pinned[C_SURV] + c->h_pinned[C_SURV] / 4;
			if (c->d_surv.need(c->surv_cap) || c->d_xs.need((size_t)c->surv_cap * 3) || c->d_cls.need(64) || c->d_res.need(c->surv_cap) || c->d_hits.need(c->surv_cap) || c->d_keys.need(c->surv_cap)) return BG_ENOMEM;
		}
		if (grow_g) c->scratch_w = c->h_pinned[C_SCRATCH] + 64;
	}
	return fail(BG_EOVERFLOW, "survivor list kept overflowing (%u entries)", c->h_pinned[C_SURV]);
}
extern "C" void *bg_batch_best_device(bg_ctx *c) { return c ? (void *)c->d_best.p : nullptr; }
extern "C" void *bg_stream(bg_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int bg_set_surv_cap(bg_ctx *c, uint32_t cap) {              // tests: start from a tiny survivor list to exercise the grow-and-redo paths
	if (!c || cap < 16) return fail(BG_EINVAL, "bg_set_surv_cap: null ctx or cap < 16");
	c->d_surv.release(); c->surv_cap = cap; c->surv_cap_forced = true;
	return BG_OK;
}
extern "C" int bg_batch_run_select(bg_ctx *c, int mode) {
	if (!c || c->kind == WORK_NONE) return fail(BG_EINVAL, "bg_batch_run: no batch uploaded");
	return run_select(c, mode);
}
extern "C" int bg_batch_run(bg_ctx *c, int mode, const uint16_t *best_in) {
	if (!c || c->kind == WORK_NONE) return fail(BG_EINVAL, "bg_batch_run: no batch uploaded");
	int rc = run_extend(c, mode, best_in);
	if (rc) return rc;
	return run_select(c, mode);
}

// Wait for the batch; if the survivor list or the generic-band scratch overflowed, grow and redo.
static int settle(bg_ctx *c) {
	if (!c->ran) return fail(BG_EINVAL, "no batch has been run");
	for (int attempt = 0; attempt < 4; ++attempt) {
		CU(cudaMemcpyAsync(c->h_pinned, c->d_counters.p, 16, cudaMemcpyDeviceToHost, c->stream));
		CU(cudaStreamSynchronize(c->stream));
		memcpy(c->h_counters, c->h_pinned, 16);
		bool grow_s = c->h_counters[C_SURV] > c->surv_cap, grow_g = c->h_counters[C_SCRATCH] > c->scratch_w;
		if (!grow_s && !grow_g) return BG_OK;
		if (grow_s) {
			c->surv_cap = c->h_counters[C_SURV] + c->h_counters[C_SURV] / 4;
			if (c->d_surv.need(c->surv_cap) || c->d_xs.need((size_t)c->surv_cap * 3) || c->d_cls.need(64) || c->d_res.need(c->surv_cap) || c->d_hits.need(c->surv_cap) || c->d_keys.need(c->surv_cap)) return BG_ENOMEM;
		}
		if (grow_g) c->scratch_w = c->h_counters[C_SCRATCH] + 64;
		int rc = run_extend(c, c->last_mode, c->have_best_in ? c->last_best_in.data() : nullptr);
		if (rc) return rc;
		rc = run_select(c, c->last_mode);
		if (rc) return rc;
	}
	return fail(BG_EOVERFLOW, "survivor list kept overflowing (%u entries)", c->h_counters[C_SURV]);
}

extern "C" int bg_batch_count(bg_ctx *c, uint64_t *nhits) {
	if (!c) return fail(BG_EINVAL, "null ctx");
	CU(cudaSetDevice(c->device));
	int rc = settle(c); if (rc) return rc;
	if (nhits) *nhits = c->h_counters[C_HITS];
	return BG_OK;
}

// hits ordered by (task, lane) on the device: radix sort of 32+4-bit keys, then a gather
static int sort_hits(bg_ctx *c, uint32_t n) {
	if (c->sorted || !n) { c->sorted = true; return BG_OK; }
	if (n > 0x7FFFFFFFu) return fail(BG_EOVERFLOW, "%u hits in one batch (the hit sort takes at most 2^31-1); use smaller batches", n);
	if (c->d_keys2.need(n) || c->d_order.need(n) || c->d_order2.need(n) || c->d_hits_sorted.need(n)) return BG_ENOMEM;
	size_t tmp = 0;
	int end_bit = 36;                                             // key = task << 4 | lane; a run list bounds the task index: sort only the bits that can differ
	if (c->kind == WORK_RUNS && c->ntasks) { end_bit = 5; while (end_bit < 36 && ((c->ntasks - 1) >> (end_bit - 4))) ++end_bit; }
	CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->d_keys.p, c->d_keys2.p, c->d_order.p, c->d_order2.p, (int)n, 0, end_bit, c->stream));
	if (c->d_sort_tmp.need(tmp + 16)) return BG_ENOMEM;
	k_iota<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_order.p, n);
	CU(cub::DeviceRadixSort::SortPairs(c->d_sort_tmp.p, tmp, c->d_keys.p, c->d_keys2.p, c->d_order.p, c->d_order2.p, (int)n, 0, end_bit, c->stream));
	k_gather_hits<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_hits.p, c->d_order2.p, n, c->d_hits_sorted.p);
	CU(cudaGetLastError());
	c->sorted = true;
	return BG_OK;
}

extern "C" int bg_batch_download(bg_ctx *c, bg_hit *hits, uint64_t cap, uint16_t *best_out) {
	if (!c) return fail(BG_EINVAL, "null ctx");
	CU(cudaSetDevice(c->device));
	int rc = settle(c); if (rc) return rc;
	uint64_t n = c->h_counters[C_HITS];
	if (hits) {
		if (cap < n) return fail(BG_EINVAL, "bg_batch_download: %llu hits, room for %llu", (unsigned long long)n, (unsigned long long)cap);
		rc = sort_hits(c, (uint32_t)n); if (rc) return rc;
		if (n) CU(cudaMemcpyAsync(hits, c->d_hits_sorted.p, n * sizeof(bg_hit), cudaMemcpyDeviceToHost, c->stream));
	}
	std::vector<uint32_t> b32;
	if (best_out) {
		b32.resize(c->nslots);
		CU(cudaMemcpyAsync(b32.data(), c->d_best.p, c->nslots * 4, cudaMemcpyDeviceToHost, c->stream));
	}
	CU(cudaStreamSynchronize(c->stream));
	if (best_out) for (uint32_t i = 0; i < c->nslots; ++i) best_out[i] = (uint16_t)std::min<uint32_t>(b32[i], 0xFFFF);
	// internal task id (run * 16 + i) -> the caller's task index
	if (hits && c->kind == WORK_TASKS) for (uint64_t i = 0; i < n; ++i) hits[i].task = c->task0[hits[i].task >> 4] + (hits[i].task & 15);
	else if (hits && c->kind == WORK_ALL) for (uint64_t i = 0; i < n; ++i) {
		const uint32_t r = hits[i].task >> 4, cl = r / c->ntiles, tile = r % c->ntiles;
		hits[i].task = cl * c->nq + tile * BG_RUN_MAX + (hits[i].task & 15);
	}
	return BG_OK;
}

extern "C" int bg_batch_stats(bg_ctx *c, bg_stats *out) {
	if (!c || !out) return fail(BG_EINVAL, "null argument");
	CU(cudaSetDevice(c->device));
	int rc = settle(c); if (rc) return rc;
	CU(cudaMemsetAsync(c->d_cells.p + 1, 0, 32, c->stream));
	if (c->nruns) k_work_stats<<<(unsigned)c->sms * 8, 256, 0, c->stream>>>(work_of(c), c->d_qi.p, c->d_clump_len.p, c->d_cells.p + 1);
	CU(cudaGetLastError());
	unsigned long long v[5] = {0, 0, 0, 0, 0};
	CU(cudaMemcpyAsync(v, c->d_cells.p, 40, cudaMemcpyDeviceToHost, c->stream));
	CU(cudaStreamSynchronize(c->stream));
	c->stats.tasks = v[1]; c->stats.nominal_cells = v[2]; c->stats.filter_cells = v[3]; c->stats.seed_steps = v[4];
	c->stats.survivors = c->h_counters[C_SURV]; c->stats.hits = c->h_counters[C_HITS]; c->stats.band_cells = v[0];
	c->stats.seed_queries = c->nseed; c->stats.seed_stride = c->SL.stride; c->stats.seed_window = c->SL.w; c->stats.seed_words = c->SL.words;
	cudaEventElapsedTime(&c->stats.ms_filter, c->ev[0], c->ev[1]);
	cudaEventElapsedTime(&c->stats.ms_extend, c->ev[1], c->ev[2]);
	cudaEventElapsedTime(&c->stats.ms_select, c->ev[2], c->ev[3]);
	*out = c->stats;
	return BG_OK;
}

__global__ void k_best16(const uint32_t *__restrict__ best, uint16_t *__restrict__ out, uint32_t n) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = (uint16_t)min(best[i], 0xFFFFu);
}

// ---------------------------------------------------------------------------------------------
// Pipelined one-call path (large run-list batches): the batch is cut into slices of consecutive runs;
// slice i+1's queries and runs travel host -> device on the copy stream while slice i's kernels
// (query prep, seed filter, banded extend) run on the compute stream.  All slices append to one survivor
// list and share the per-slot minima, so the selection (which needs the batch-wide minimum of a read and
// its reverse complement, burst.c:4229/4497) runs once at the end, followed by the hit sort and the
// device -> host copy.  Results are identical to the single-batch path.
// ---------------------------------------------------------------------------------------------
static int align_pipelined(bg_ctx *c, const bg_queries *Q, const bg_run *runs, uint64_t nruns, int mode,
		uint16_t *best_inout, bg_hit *hits, uint64_t cap, uint64_t *nhits) {
	CU(cudaSetDevice(c->device));
	const uint32_t nq = Q->nq;
	c->kind = WORK_NONE; c->ran = false;
	// ---- window layout from a sample of the batch (the device re-checks every query against it) ----
	SeedLayout SL = {0, 0, 0, 0, 0, 0, 0}; uint32_t npmax = 1;
	{
		uint32_t hist[32]; memset(hist, 0, sizeof(hist));
		const uint32_t step = std::max<uint32_t>(1, nq / 8192); uint32_t ns = 0; uint64_t mxlen = 0;
		for (uint32_t q = 0; q < nq; q += step, ++ns) {
			const uint64_t len = Q->offset[q + 1] - Q->offset[q]; const uint32_t np = Q->budget[q] + 1u;
			if (np <= SEED_NP_MAX) ++hist[std::min<uint64_t>(len / np, 31)];
			mxlen = std::max(mxlen, len);
		}
		c->mstage = stage_len(mxlen); c->wide_possible = long_queries(mxlen); c->len_hint = (uint32_t)mxlen;
		SL = choose_layout(c, hist, ns);
		if (SL.stride) {
			uint64_t sum = 0, cnt = 0; uint32_t mx = 1;
			for (uint32_t q = 0; q < nq; q += step) {
				const uint64_t len = Q->offset[q + 1] - Q->offset[q]; const uint32_t np = Q->budget[q] + 1u;
				if (np <= SL.np_max && len / np >= SL.w + SL.stride - 1) { sum += np; ++cnt; mx = std::max(mx, np); }
			}
			seed_sizes(c, SL, cnt ? (uint32_t)((sum + cnt - 1) / cnt) : 1, mx, npmax);
			SL.np_max = npmax;                                   // queries with more stretches than the sample showed go to k_filter
		}
	}
	// ---- batch-wide device state ----
	const uint64_t ntasks = nruns * BG_RUN_MAX;
	if (!c->surv_cap) c->surv_cap = 1u << 20;
	uint64_t want = std::min<uint64_t>(ntasks * 16, std::max<uint64_t>(c->surv_cap, 4ull * nq + ntasks / 8));
	if (want > c->surv_cap || !c->d_surv.p) c->surv_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, 1024), 0xFFFFFFF0ull);
	const int nsl = (int)std::min<uint64_t>((uint64_t)c->pipe_slices, std::max<uint64_t>(1, nruns / (uint64_t)c->pipe_min_runs));
	// slice boundaries: sizes shrink geometrically so that little work is left when the last copy lands
	std::vector<uint64_t> cut(nsl + 1, 0);
	{
		// growing slices when the kernels are the longer leg (nibble-packed reads: first copy short, copies hide behind the kernels),
		// equal slices when the copies are (one byte per base)
		const double ratio = (c->pipe_ratio ? c->pipe_ratio : ((Q->flags & BG_Q_PACKED4) ? 140 : 100)) / 100.0; double tot = 0, w = 1, acc = 0;
		for (int i = 0; i < nsl; ++i, w *= ratio) tot += w;
		w = 1;
		for (int i = 0; i < nsl; ++i, w *= ratio) { acc += w; cut[i + 1] = std::min<uint64_t>(nruns, (uint64_t)(nruns * (acc / tot) + 0.5)); }
		cut[nsl] = nruns;
	}
	const bool dbg = getenv("BURST_B200_TIMING") != nullptr;
	for (int attempt = 0; attempt < 4; ++attempt) {
		if (c->d_surv.need(c->surv_cap) || c->d_xs.need((size_t)c->surv_cap * 3) || c->d_cls.need(64) || c->d_res.need(c->surv_cap) || c->d_hits.need(c->surv_cap) || c->d_keys.need(c->surv_cap)) return BG_ENOMEM;
			if (c->d_best.need(Q->nslots) || c->d_best16.need(Q->nslots) || c->d_counters.need(64) || c->d_cells.need(8) || c->d_first.need(128)) return BG_ENOMEM;
		cudaStream_t cs = c->stream, ps = c->copy_stream;
		if (best_inout) CU(cudaMemcpyAsync(c->d_best16.p, best_inout, (size_t)Q->nslots * 2, cudaMemcpyHostToDevice, cs));
		k_init_best<<<(Q->nslots + 255) / 256, 256, 0, cs>>>(c->d_best.p, best_inout ? c->d_best16.p : nullptr, Q->nslots);
		CU(cudaMemsetAsync(c->d_counters.p, 0, 256, cs));
		CU(cudaMemsetAsync(c->d_cells.p, 0, 8, cs));
		CU(cudaMemsetAsync(c->d_first.p, 0, 128 * 4, cs));        // [0,64) first survivor of each slice, [64,128) its queries left to k_filter
		CU(cudaEventRecord(c->ev[0], cs));
		CU(cudaStreamWaitEvent(ps, c->ev[0], 0));                 // copies of this call start after earlier work of the context
		for (int i = 0; i < nsl; ++i) {
			const uint64_t ra = cut[i], rb = cut[i + 1];
			if (ra >= rb) continue;
			// queries this slice touches
			uint32_t qa = 0xFFFFFFFFu, qb = 0;
			for (uint64_t r = ra; r < rb; ++r) {
				const uint32_t a = runs[r].query0, n = runs[r].nq;
				if (!n || n > BG_RUN_MAX || (uint64_t)a + n > nq) { cudaStreamSynchronize(c->copy_stream); cudaStreamSynchronize(c->stream); }
				if (!n || n > BG_RUN_MAX || (uint64_t)a + n > nq)
					return fail(BG_EINVAL, "bg_align_runs: run %llu is malformed (nq must be 1..%d and query0+nq within the batch)", (unsigned long long)r, BG_RUN_MAX);
				qa = std::min(qa, a); qb = std::max(qb, a + n);
			}
			const uint32_t n = qb - qa; const uint64_t base = Q->offset[qa], bytes = Q->offset[qb] - base;
			if (Q->offset[qb] < base) { cudaStreamSynchronize(c->copy_stream); cudaStreamSynchronize(c->stream); return fail(BG_EINVAL, "bg_align_runs: query offsets are not ascending"); }
			Slice &S = c->sl[i % NSLICEBUF];
			if (S.qoff.need((size_t)n + 1) || S.budget.need(n) || S.slot.need(n) || S.qi.need(n) ||
			    S.qnib.need(bytes / 8 + 3ull * n + 8) || S.runs.need(rb - ra + 1)) { cudaStreamSynchronize(ps); cudaStreamSynchronize(cs); return BG_ENOMEM; }   // earlier slices still read the caller's arrays
			if (i >= NSLICEBUF) CU(cudaStreamWaitEvent(ps, S.computed, 0));   // the slice that used these buffers last is done with them
			{ int rc = copy_codes(Q, base, base + bytes, S.packed, S.codes, ps, ps, S.copied); if (rc) { cudaStreamSynchronize(ps); cudaStreamSynchronize(cs); return rc; } }
			CU(cudaMemcpyAsync(S.qoff.p, Q->offset + qa, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, ps));
			CU(cudaMemcpyAsync(S.budget.p, Q->budget + qa, (size_t)n * 2, cudaMemcpyHostToDevice, ps));
			CU(cudaMemcpyAsync(S.slot.p, Q->slot + qa, (size_t)n * 4, cudaMemcpyHostToDevice, ps));
			CU(cudaMemcpyAsync(S.runs.p, runs + ra, (rb - ra) * sizeof(bg_run), cudaMemcpyHostToDevice, ps));
			CU(cudaEventRecord(S.copied, ps));
			CU(cudaStreamWaitEvent(cs, S.copied, 0));
			k_qinfo<<<(n + 255) / 256, 256, 0, cs>>>(S.qoff.p, S.budget.p, S.slot.p, n, Q->nslots, S.qi.p, nullptr, c->d_counters.p, base);
			k_qprep<<<(n + 127) / 128, 128, 0, cs>>>(S.codes.p, S.qi.p, n, SL, S.qnib.p, c->d_counters.p + 9, c->d_first.p + 64 + i);
			k_check_runs<<<(unsigned)((rb - ra + 255) / 256), 256, 0, cs>>>(S.runs.p, rb - ra, qa, n, c->d_counters.p);
			CU(cudaMemcpyAsync(c->d_first.p + i, c->d_counters.p + C_SURV, 4, cudaMemcpyDeviceToDevice, cs));
			BatchDev B; B.codes = S.codes.p; B.qnib = S.qnib.p; B.qi = S.qi.p;
			B.W.runs = S.runs.p; B.W.nruns = rb - ra; B.W.nq = n; B.W.ntiles = 0; B.W.first_clump = c->first_clump; B.W.num_clumps = c->num_clumps;
			B.W.q_base = qa; B.W.run_base = (uint32_t)ra;
			int rc = launch_filters(c, cs, B, SL, npmax, SL.stride != 0, true, c->d_first.p + 64 + i); if (rc) return rc;
			rc = launch_extend(c, cs, B, mode, c->d_first.p + i); if (rc) return rc;
			CU(cudaEventRecord(S.computed, cs));
			if (dbg && rb == nruns) { CU(cudaEventRecord(c->ev[1], ps)); CU(cudaEventRecord(c->ev[2], cs)); }
		}
		k_select<<<(unsigned)c->sms * 4, 256, 0, cs>>>(c->d_surv.p, c->d_res.p, c->d_best.p, c->d_counters.p, c->surv_cap, c->d_hits.p, c->d_keys.p, mode);
		CU(cudaGetLastError());
		CU(cudaMemcpyAsync(c->h_pinned, c->d_counters.p, 16, cudaMemcpyDeviceToHost, cs));
		CU(cudaStreamSynchronize(cs));
		CU(cudaStreamSynchronize(ps));
		memcpy(c->h_counters, c->h_pinned, 16);
		if (c->h_counters[C_ERR]) return fail(BG_EINVAL, "bg_align_runs: a query or run of the batch is malformed (query lengths >= 1, budgets <= 254 (burst.c:3076), slots < nslots, runs within the batch)");
		const bool grow_s = c->h_counters[C_SURV] > c->surv_cap, grow_g = c->h_counters[C_SCRATCH] > c->scratch_w;
		if (!grow_s && !grow_g) {
			const uint64_t n = c->h_counters[C_HITS];
			if (n > cap) { *nhits = n; return fail(BG_EOVERFLOW, "bg_align_runs_into: %llu hits, room for %llu", (unsigned long long)n, (unsigned long long)cap); }
			c->sorted = false;
			int rc = sort_hits(c, (uint32_t)n); if (rc) return rc;
			if (n) CU(cudaMemcpyAsync(hits, c->d_hits_sorted.p, n * sizeof(bg_hit), cudaMemcpyDeviceToHost, cs));
			if (best_inout) {
				k_best16<<<(Q->nslots + 255) / 256, 256, 0, cs>>>(c->d_best.p, c->d_best16.p, Q->nslots);
				CU(cudaMemcpyAsync(best_inout, c->d_best16.p, (size_t)Q->nslots * 2, cudaMemcpyDeviceToHost, cs));
			}
			if (dbg) CU(cudaEventRecord(c->ev[3], cs));
			CU(cudaStreamSynchronize(cs));
			if (dbg) {
				float a = 0, b = 0, d = 0; cudaEventElapsedTime(&a, c->ev[0], c->ev[1]); cudaEventElapsedTime(&b, c->ev[0], c->ev[2]); cudaEventElapsedTime(&d, c->ev[0], c->ev[3]);
				fprintf(stderr, "[burst_b200] one-call timing: last copy done %.3f ms, last slice computed %.3f ms, hits on host %.3f ms (%d slices)\n", a, b, d, nsl);
			}
			*nhits = n;
			return BG_OK;
		}
		if (grow_s) c->surv_cap = c->h_counters[C_SURV] + c->h_counters[C_SURV] / 4;
		if (grow_g) c->scratch_w = c->h_counters[C_SCRATCH] + 64;
	}
	return fail(BG_EOVERFLOW, "survivor list kept overflowing (%u entries)", c->h_counters[C_SURV]);
}

static bool pipeline_applies(const bg_ctx *c, const bg_queries *Q, uint64_t nruns) {
	return c->pipe_slices >= 2 && Q && Q->nq && nruns >= 2ull * (uint64_t)c->pipe_min_runs && nruns < (1ull << 28) && c->num_clumps;
}

extern "C" int bg_align_runs_into(bg_ctx *c, const bg_queries *Q, const bg_run *runs, uint64_t nruns, int mode,
		uint16_t *best_inout, bg_hit *hits, uint64_t cap, uint64_t *nhits) {
	if (!c || !Q || !runs || !nhits || (!hits && cap)) return fail(BG_EINVAL, "bg_align_runs_into: null argument");
	if (pipeline_applies(c, Q, nruns)) {
		if (!Q->codes || !Q->offset || !Q->budget || !Q->slot) return fail(BG_EINVAL, "bg_align_runs_into: null query array");
		if (Q->flags & ~(uint32_t)BG_Q_PACKED4) return fail(BG_EINVAL, "bg_align_runs_into: unknown query flags %u", Q->flags);
		return align_pipelined(c, Q, runs, nruns, mode, best_inout, hits, cap, nhits);
	}
	int rc = bg_batch_upload_runs(c, Q, runs, nruns); if (rc) return rc;
	rc = bg_batch_run(c, mode, best_inout); if (rc) return rc;
	uint64_t n = 0;
	rc = bg_batch_count(c, &n); if (rc) return rc;
	*nhits = n;
	if (n > cap) return fail(BG_EOVERFLOW, "bg_align_runs_into: %llu hits, room for %llu", (unsigned long long)n, (unsigned long long)cap);
	return bg_batch_download(c, hits, cap, best_inout);
}

extern "C" int bg_align_bunches_into(bg_ctx *c, const bg_reads *R, uint32_t qbunch, const uint32_t *cand_off, const uint32_t *cand, uint32_t nbunch,
		int mode, uint16_t *best_inout, bg_hit *hits, uint64_t cap, uint64_t *nhits) {
	if (!c || !R || !cand_off || !cand || !nhits || (!hits && cap)) return fail(BG_EINVAL, "bg_align_bunches_into: null argument");
	if (!c->num_clumps) return fail(BG_EINVAL, "bg_align_bunches_into: no database loaded");
	if (!R->reads || !R->len || !R->budget || !R->strand || !R->nreads || !R->nq) return fail(BG_EINVAL, "bg_align_bunches_into: null or empty read arrays");
	if (R->flags != BG_R_PACKED4 && R->flags != BG_R_PACKED2) return fail(BG_EINVAL, "bg_align_bunches_into: flags must be BG_R_PACKED4 or BG_R_PACKED2");
	if (!qbunch || qbunch > BG_RUN_MAX) return fail(BG_EINVAL, "bg_align_bunches_into: bunch size %u (must be 1..%d)", qbunch, BG_RUN_MAX);
	if ((uint64_t)nbunch * qbunch < R->nq || (uint64_t)(nbunch - 1) * qbunch >= R->nq) return fail(BG_EINVAL, "bg_align_bunches_into: %u bunches of %u do not cover %u strands", nbunch, qbunch, R->nq);
	const uint64_t nruns = cand_off[nbunch];
	if (cand_off[0] != 0 || nruns >= (1ull << 28)) return fail(BG_EINVAL, "bg_align_bunches_into: candidate offsets must start at 0 and end below 2^28");
	const auto wall0 = std::chrono::steady_clock::now();
	CU(cudaSetDevice(c->device));
	c->kind = WORK_NONE; c->ran = false;
	const uint32_t nq = R->nq, nr = R->nreads;
	cudaStream_t cs = c->stream, ps = c->copy_stream;
	// ---- the small arrays leave first, so that they travel while the host looks at the lengths (single-slice calls, first attempt) ----
	static const bool no_prep2 = getenv("BURST_B200_NO_PREP2") != nullptr;   // debugging: the two-kernel preparation also for 2-bit reads
	static const int want_slices = getenv("BURST_B200_COMPACT_SLICES") ? std::max(1, std::min(32, atoi(getenv("BURST_B200_COMPACT_SLICES")))) : 1;
	const int nsl = (int)std::min<uint64_t>((uint64_t)want_slices, std::max<uint64_t>(1, nruns / (uint64_t)c->pipe_min_runs));
	bool early = nsl == 1;
	if (early) {
		if (c->d_rlen.need(nr) || c->d_rbud.need(nr) || c->d_strand.need(nq) || c->d_candoff.need((size_t)nbunch + 1) || c->d_cand.need(nruns + 1)) return BG_ENOMEM;
		CU(cudaMemcpyAsync(c->d_rlen.p, R->len, (size_t)nr * 2, cudaMemcpyHostToDevice, ps));
		CU(cudaMemcpyAsync(c->d_rbud.p, R->budget, (size_t)nr * 2, cudaMemcpyHostToDevice, ps));
		CU(cudaMemcpyAsync(c->d_candoff.p, cand_off, ((size_t)nbunch + 1) * 4, cudaMemcpyHostToDevice, ps));
		CU(cudaEventRecord(c->sl[1].copied, ps));
		CU(cudaMemcpyAsync(c->d_strand.p, R->strand, (size_t)nq * 4, cudaMemcpyHostToDevice, ps));
		if (nruns) CU(cudaMemcpyAsync(c->d_cand.p, cand, (size_t)nruns * 4, cudaMemcpyHostToDevice, ps));
		CU(cudaEventRecord(c->sl[2].computed, ps));                  // "strands and candidates are on the device"
	}
	// ---- one pass over the read lengths on the host: bytes of the packed stream, the longest read (bounds every device buffer, so that
	//      nothing has to come back from the device before the end), and a sample for the window layout (as the run-list path does) ----
	// (sum and bitwise OR vectorise; the OR of all lengths is an upper bound of the longest, at most twice it: it only sizes buffers)
	uint64_t total = 0; uint32_t maxlen = 0;
	for (uint32_t r0 = 0; r0 < nr; r0 += 32768) {
		const uint32_t r1 = std::min(nr, r0 + 32768u); uint32_t sum = 0; uint16_t orv = 0;
		for (uint32_t r = r0; r < r1; ++r) { sum += R->len[r]; orv |= R->len[r]; }
		total += sum; maxlen |= orv;
	}
	const uint64_t rbytes = R->flags == BG_R_PACKED2 ? (total + 3) / 4 : (total + 1) / 2, ncodes_max = (uint64_t)nq * (((uint64_t)maxlen + 15) & ~15ull);
	SeedLayout SL = {0, 0, 0, 0, 0, 0, 0}; uint32_t npmax = 1;
	{
		uint32_t hist[32]; memset(hist, 0, sizeof(hist));
		const uint32_t step = std::max<uint32_t>(1, nr / 8192); uint32_t ns = 0;
		for (uint32_t r = 0; r < nr; r += step, ++ns) { const uint32_t np = R->budget[r] + 1u; if (np <= SEED_NP_MAX) ++hist[std::min<uint32_t>(R->len[r] / np, 31)]; }
		SL = choose_layout(c, hist, ns);
		if (SL.stride) {
			uint64_t sum = 0, cnt = 0; uint32_t mx = 1;
			for (uint32_t r = 0; r < nr; r += step) { const uint32_t np = R->budget[r] + 1u; if (np <= SL.np_max && R->len[r] / np >= SL.w + SL.stride - 1) { sum += np; ++cnt; mx = std::max(mx, np); } }
			seed_sizes(c, SL, cnt ? (uint32_t)((sum + cnt - 1) / cnt) : 1, mx, npmax);
			SL.np_max = npmax;                                       // reads with more stretches than the sample showed go to k_filter
		}
		{ uint32_t smax = 1; for (uint32_t r = 0; r < nr; r += step) smax = std::max<uint32_t>(smax, R->len[r]); c->mstage = stage_len(smax); c->wide_possible = long_queries(smax); c->len_hint = smax; }   // (a longer read still works: it reads global memory)
	}
	if (c->d_rlen.need(nr) || c->d_rbud.need(nr) || c->d_strand.need(nq) || c->d_candoff.need((size_t)nbunch + 1) || c->d_runs.need(nruns + 1) || c->d_cand.need(nruns + 1) ||
	    c->d_rl64.need((size_t)nr + 1) || c->d_sl64.need((size_t)nq + 1) || c->d_roff.need((size_t)nr + 1) || c->d_qoff.need((size_t)nq + 1) ||
	    c->d_qi.need(nq) || c->d_best.need(nr) || c->d_best16.need(nr) || c->d_counters.need(64) || c->d_cells.need(8) || c->d_first.need(128) ||
	    c->d_packed.need(rbytes + 32) || c->d_codes.need(ncodes_max + 32) || c->d_qnib.need(ncodes_max / 8 + 3ull * nq + 8)) return BG_ENOMEM;
	size_t tmp1 = 0, tmp2 = 0;
	CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp1, c->d_rl64.p, c->d_roff.p, (int)nr + 1, cs));
	CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, c->d_sl64.p, (unsigned long long *)c->d_qoff.p, (int)nq + 1, cs));
	if (c->d_sort_tmp.need(std::max(tmp1, tmp2) + 16)) return BG_ENOMEM;
	const uint64_t ntasks = nruns * BG_RUN_MAX;
	if (!c->surv_cap) c->surv_cap = 1u << 20;
	{ uint64_t want = std::min<uint64_t>(ntasks * 16, std::max<uint64_t>(c->surv_cap, 4ull * nq + ntasks / 8)); want = std::max<uint64_t>(want, 1024);
	  if (c->surv_cap_forced) { want = c->surv_cap; c->surv_cap_forced = false; }
	  if (want > c->surv_cap || !c->d_surv.p) c->surv_cap = (uint32_t)std::min<uint64_t>(want, 0xFFFFFFF0ull); }
	for (int attempt = 0; attempt < 4; ++attempt) {
		if (c->d_surv.need(c->surv_cap) || c->d_xs.need((size_t)c->surv_cap * 3) || c->d_cls.need(64) || c->d_res.need(c->surv_cap) || c->d_hits.need(c->surv_cap) || c->d_keys.need(c->surv_cap)) return BG_ENOMEM;
			// ---- host -> device on the copy stream: the reads and their lengths/budgets first, then the strands and candidates slice by slice;
		//      the kernels of slice i (strand records, codes, runs, seed filter, banded sweep) run while slice i+1 travels ----
		static cudaEvent_t te0 = nullptr;
		const bool dbg = getenv("BURST_B200_TIMING") != nullptr;
		static cudaEvent_t te[4] = {nullptr, nullptr, nullptr, nullptr};
		if (dbg && !te0) { cudaEventCreate(&te0); for (int i = 1; i < 4; ++i) cudaEventCreate(&te[i]); }
		te[0] = te0;
		const auto wall1 = std::chrono::steady_clock::now();
		if (dbg) cudaEventRecord(te0, cs);
		// contexts over one shared database: this batch's filter + sweep behind those of the batch queued last (taken below, before the first
		// of them is launched; the copies and the light preparation kernels do not wait and run in the shadow of the other batch)
		std::unique_lock<std::mutex> gate_lock;
		if (best_inout) CU(cudaMemcpyAsync(c->d_best16.p, best_inout, (size_t)nr * 2, cudaMemcpyHostToDevice, cs));
		k_init_best<<<(nr + 255) / 256, 256, 0, cs>>>(c->d_best.p, best_inout ? c->d_best16.p : nullptr, nr);
		CU(cudaMemsetAsync(c->d_counters.p, 0, 256, cs));
		CU(cudaMemsetAsync(c->d_cells.p, 0, 8, cs));
		CU(cudaMemsetAsync(c->d_first.p, 0, 128 * 4, cs));
		CU(cudaEventRecord(c->ev[0], cs));
		if (!early) {
			CU(cudaStreamWaitEvent(ps, c->ev[0], 0));               // (the previous attempt is done with the buffers)
			CU(cudaMemcpyAsync(c->d_rlen.p, R->len, (size_t)nr * 2, cudaMemcpyHostToDevice, ps));
			CU(cudaMemcpyAsync(c->d_rbud.p, R->budget, (size_t)nr * 2, cudaMemcpyHostToDevice, ps));
			CU(cudaMemcpyAsync(c->d_candoff.p, cand_off, ((size_t)nbunch + 1) * 4, cudaMemcpyHostToDevice, ps));
			CU(cudaEventRecord(c->sl[1].copied, ps));
		}
		CU(cudaMemcpyAsync(c->d_packed.p, R->reads, rbytes, cudaMemcpyHostToDevice, ps));
		CU(cudaEventRecord(c->sl[0].copied, ps));
		CU(cudaStreamWaitEvent(cs, c->sl[1].copied, 0));
		k_compact_rlen<<<(nr + 256) / 256, 256, 0, cs>>>(c->d_rlen.p, nr, c->d_rl64.p);
		CU(cub::DeviceScan::ExclusiveSum(c->d_sort_tmp.p, tmp1, c->d_rl64.p, c->d_roff.p, (int)nr + 1, cs));
		c->nq = nq; c->nslots = nr; c->SL = SL; c->seed_npmax = npmax;
		c->kind = WORK_RUNS; c->nruns = nruns; c->ntasks = ntasks; c->ntiles = 0;
		// slices of whole bunches, growing (the first one short so that the kernels start early)
		// (measured, profiles/: one slice is fastest on the bench workload -- the per-slice launches and kernel tails cost more than the
		//  ~0.8 ms of copies they hide; BURST_B200_COMPACT_SLICES cuts the batch for hosts with slower links)
		std::vector<uint32_t> cut(nsl + 1, 0);
		{ double tot = 0, w = 1, acc = 0; const double ratio = 1.4;
		  for (int i = 0; i < nsl; ++i, w *= ratio) tot += w;
		  w = 1;
		  for (int i = 0; i < nsl; ++i, w *= ratio) { acc += w; cut[i + 1] = (uint32_t)std::min<uint64_t>(nbunch, (uint64_t)(nbunch * (acc / tot) + 0.5)); }
		  cut[nsl] = nbunch; }
		int rc = BG_OK;
		for (int i = 0; i < nsl; ++i) {
			const uint32_t ba = cut[i], bb = cut[i + 1];
			if (ba >= bb) continue;
			const uint32_t qa = (uint32_t)std::min<uint64_t>(nq, (uint64_t)ba * qbunch), qb = (uint32_t)std::min<uint64_t>(nq, (uint64_t)bb * qbunch), n = qb - qa;
			const uint32_t ra = cand_off[ba], rb = cand_off[bb];
			if (rb < ra || rb > nruns) { cudaStreamSynchronize(ps); cudaStreamSynchronize(cs); return fail(BG_EINVAL, "bg_align_bunches_into: candidate offsets are not ascending"); }
			if (!n) continue;
			Slice &S = c->sl[i % NSLICEBUF];
			const uint64_t scodes = (uint64_t)n * (((uint64_t)maxlen + 15) & ~15ull);
			if (S.sl64.need((size_t)n + 1) || S.qoff.need((size_t)n + 1) || S.qi.need(n) || S.codes.need(scodes + 32) || S.qnib.need(scodes / 8 + 3ull * n + 8) || S.runs.need((size_t)(rb - ra) + 1)) {
				cudaStreamSynchronize(ps); cudaStreamSynchronize(cs); return BG_ENOMEM;
			}
			if (!early) {
				CU(cudaMemcpyAsync(c->d_strand.p + qa, R->strand + qa, (size_t)n * 4, cudaMemcpyHostToDevice, ps));
				if (rb > ra) CU(cudaMemcpyAsync(c->d_cand.p + ra, cand + ra, (size_t)(rb - ra) * 4, cudaMemcpyHostToDevice, ps));
				CU(cudaEventRecord(S.computed, ps));                   // (used here as "slice i is on the device")
				CU(cudaStreamWaitEvent(cs, S.computed, 0));
			} else CU(cudaStreamWaitEvent(cs, c->sl[2].computed, 0));
			k_compact_slen<<<(n + 256) / 256, 256, 0, cs>>>(c->d_rlen.p, nr, c->d_strand.p + qa, n, S.sl64.p, c->d_counters.p);
			CU(cub::DeviceScan::ExclusiveSum(c->d_sort_tmp.p, tmp2, S.sl64.p, (unsigned long long *)S.qoff.p, (int)n + 1, cs));
			k_compact_qinfo<<<(n + 255) / 256, 256, 0, cs>>>((const unsigned long long *)S.qoff.p, c->d_rlen.p, c->d_rbud.p, c->d_strand.p + qa, n, nr, S.qi.p, c->d_counters.p + 16, c->d_counters.p);
			if (rb > ra) k_compact_runs<<<(unsigned)((rb - ra + 255) / 256), 256, 0, cs>>>(c->d_candoff.p, c->d_cand.p, nbunch, ra, rb - ra, qbunch, nq, S.runs.p, c->d_counters.p);
			if (i == 0) CU(cudaStreamWaitEvent(cs, c->sl[0].copied, 0));   // the packed reads
			if (R->flags == BG_R_PACKED2 && !no_prep2)
				k_compact_prep2<<<(unsigned)std::min<uint64_t>(((uint64_t)n * 8 + 255) / 256, (uint64_t)c->sms * 8), 256, 0, cs>>>((const uint32_t *)c->d_packed.p, c->d_roff.p, c->d_rlen.p, c->d_rbud.p, c->d_strand.p + qa, (const unsigned long long *)S.qoff.p, n, nr,
					S.codes.p, S.qi.p, SL, S.qnib.p, c->d_counters.p + 9, c->d_first.p + 64 + i);
			else {
				k_compact_codes<<<(unsigned)(((uint64_t)n * 8 + 255) / 256), 256, 0, cs>>>(c->d_packed.p, R->flags, c->d_roff.p, c->d_rlen.p, c->d_strand.p + qa, (const unsigned long long *)S.qoff.p, n, nr, S.codes.p);
				k_qprep<<<(n + 127) / 128, 128, 0, cs>>>(S.codes.p, S.qi.p, n, SL, S.qnib.p, c->d_counters.p + 9, c->d_first.p + 64 + i);
			}
			CU(cudaMemcpyAsync(c->d_first.p + i, c->d_counters.p + C_SURV, 4, cudaMemcpyDeviceToDevice, cs));
			CU(cudaGetLastError());
			BatchDev B; B.codes = S.codes.p; B.qnib = S.qnib.p; B.qi = S.qi.p;
			B.W.runs = S.runs.p; B.W.nruns = rb - ra; B.W.nq = n; B.W.ntiles = 0; B.W.first_clump = c->first_clump; B.W.num_clumps = c->num_clumps;
			B.W.q_base = qa; B.W.run_base = ra;
			if (c->gate && !gate_lock.owns_lock()) { gate_lock = std::unique_lock<std::mutex>(c->gate->m); if (c->gate->last && c->gate->last != c->gate_done) CU(cudaStreamWaitEvent(cs, c->gate->last, 0)); }
			if (i == 0) CU(cudaEventRecord(c->ev[0], cs));
			rc = launch_filters(c, cs, B, SL, npmax, SL.stride != 0, true, c->d_first.p + 64 + i); if (rc) return rc;
			rc = launch_extend(c, cs, B, mode, c->d_first.p + i); if (rc) return rc;
		}
		CU(cudaEventRecord(c->ev[1], cs)); CU(cudaEventRecord(c->ev[2], cs));
		k_select<<<(unsigned)c->sms * 4, 256, 0, cs>>>(c->d_surv.p, c->d_res.p, c->d_best.p, c->d_counters.p, c->surv_cap, c->d_hits.p, c->d_keys.p, mode);
		CU(cudaGetLastError());
		CU(cudaEventRecord(c->ev[3], cs));
		if (c->gate && gate_lock.owns_lock()) { CU(cudaEventRecord(c->gate_done, cs)); c->gate->last = c->gate_done; gate_lock.unlock(); }
		CU(cudaMemcpyAsync(c->h_pinned, c->d_counters.p, 64, cudaMemcpyDeviceToHost, cs));
		CU(cudaStreamSynchronize(cs));
		memcpy(c->h_counters, c->h_pinned, 16);
		c->nseed = c->h_pinned[9]; c->last_mode = mode; c->have_best_in = false; c->ran = true; c->sorted = false;
		if (c->h_counters[C_ERR]) { c->kind = WORK_NONE; return fail(BG_EINVAL, "bg_align_bunches_into: malformed batch (strands must name reads < nreads, read lengths >= 1, budgets <= 254 (burst.c:3076), ascending candidate offsets)"); }
		const bool grow_s = c->h_counters[C_SURV] > c->surv_cap, grow_g = c->h_counters[C_SCRATCH] > c->scratch_w;
		if (!grow_s && !grow_g) {
			const uint64_t n = c->h_counters[C_HITS];
			*nhits = n;
			if (n > cap) return fail(BG_EOVERFLOW, "bg_align_bunches_into: %llu hits, room for %llu", (unsigned long long)n, (unsigned long long)cap);
			if (dbg) cudaEventRecord(te[1], cs);
			rc = sort_hits(c, (uint32_t)n); if (rc) return rc;
			if (dbg) cudaEventRecord(te[2], cs);
			if (n) CU(cudaMemcpyAsync(hits, c->d_hits_sorted.p, n * sizeof(bg_hit), cudaMemcpyDeviceToHost, cs));
			if (best_inout) {
				k_best16<<<(nr + 255) / 256, 256, 0, cs>>>(c->d_best.p, c->d_best16.p, nr);
				CU(cudaMemcpyAsync(best_inout, c->d_best16.p, (size_t)nr * 2, cudaMemcpyDeviceToHost, cs));
			}
			if (dbg) cudaEventRecord(te[3], cs);
			CU(cudaStreamSynchronize(cs));
			if (dbg) {
				float a = 0, b = 0, d = 0, e = 0, f = 0, g = 0; 
				cudaEventElapsedTime(&a, te[0], c->ev[0]); cudaEventElapsedTime(&b, c->ev[0], c->ev[1]); cudaEventElapsedTime(&d, c->ev[1], c->ev[2]); cudaEventElapsedTime(&e, c->ev[2], te[1]);
				cudaEventElapsedTime(&f, te[1], te[2]); cudaEventElapsedTime(&g, te[2], te[3]);
				(void)d;
				const auto wall2 = std::chrono::steady_clock::now();
				fprintf(stderr, "[burst_b200] compact call: until the first filter launch %.3f ms, all slices (prep + filter + extend, copies behind them) %.3f, select + counters %.3f, sort %.3f, D2H %.3f (host: %.3f ms before the first enqueue, %.3f ms in all)\n", a, b, e, f, g,
					std::chrono::duration<double, std::milli>(wall1 - wall0).count(), std::chrono::duration<double, std::milli>(wall2 - wall0).count());
			}
			c->kind = WORK_NONE; c->ran = false;                    // the per-slice buffers are not a resident batch: nothing to re-run or to take statistics of
			return BG_OK;
		}
		if (grow_s) c->surv_cap = c->h_counters[C_SURV] + c->h_counters[C_SURV] / 4;
		if (grow_g) c->scratch_w = c->h_counters[C_SCRATCH] + 64;
		early = false;
	}
	return fail(BG_EOVERFLOW, "survivor list kept overflowing (%u entries)", c->h_counters[C_SURV]);
}

extern "C" int bg_load_acx(bg_ctx *c, const uint32_t *lens, const uint8_t *postings, uint64_t post_bytes, int word_len, int big, const uint32_t *bad, uint32_t nbad) {
	if (!c || !lens || (!postings && post_bytes) || (!bad && nbad)) return fail(BG_EINVAL, "bg_load_acx: null argument");
	if (word_len != 12 && word_len != 15) return fail(BG_EINVAL, "bg_load_acx: word length %d (must be 12 or 15)", word_len);
	if (!c->num_clumps) return fail(BG_EINVAL, "bg_load_acx: load the database first");
	CU(cudaSetDevice(c->device));
	const unsigned long long nk = 1ull << (2 * word_len);
	c->acx_n = 0;
	if (c->d_acx_off.need(nk + 1) || c->d_post.need(post_bytes + 32) || c->d_bad.need((size_t)nbad + 1)) return BG_ENOMEM;
	{	// lengths -> byte sizes -> offsets, through a bounded staging buffer
		const unsigned long long SLAB = 64ull << 20;                   // entries per piece
		DBuf<uint32_t> stage; if (stage.need(std::min(nk, SLAB))) return BG_ENOMEM;
		for (unsigned long long a = 0; a < nk; a += SLAB) {
			const unsigned long long n = std::min(SLAB, nk - a);
			CU(cudaMemcpyAsync(stage.p, lens + a, n * 4, cudaMemcpyHostToDevice, c->stream));
			k_acx_sizes<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(stage.p, n, big, c->d_acx_off.p + a);   // (writes n + 1 entries; the last is overwritten by the next piece)
			CU(cudaStreamSynchronize(c->stream));
		}
		CU(cudaMemsetAsync(c->d_acx_off.p + nk, 0, 8, c->stream));
		size_t tmp = 0;
		CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->d_acx_off.p, c->d_acx_off.p, (int)(nk + 1), c->stream));
		if (c->d_sort_tmp.need(tmp + 16)) { stage.release(); return BG_ENOMEM; }
		CU(cub::DeviceScan::ExclusiveSum(c->d_sort_tmp.p, tmp, c->d_acx_off.p, c->d_acx_off.p, (int)(nk + 1), c->stream));
		unsigned long long total = 0;
		CU(cudaMemcpyAsync(&total, c->d_acx_off.p + nk, 8, cudaMemcpyDeviceToHost, c->stream));
		CU(cudaStreamSynchronize(c->stream));
		stage.release();
		if (total != post_bytes) return fail(BG_EINVAL, "bg_load_acx: the lengths describe %llu bytes of postings, %llu given", total, (unsigned long long)post_bytes);
	}
	if (post_bytes) CU(cudaMemcpyAsync(c->d_post.p, postings, post_bytes, cudaMemcpyHostToDevice, c->stream));
	CU(cudaMemsetAsync(c->d_post.p + post_bytes, 0, 32, c->stream));
	if (nbad) CU(cudaMemcpyAsync(c->d_bad.p, bad, (size_t)nbad * 4, cudaMemcpyHostToDevice, c->stream));
	{ int rc = acx_scratch(c); if (rc) return rc; }
	c->acx_n = word_len; c->acx_big = big; c->acx_nbad = nbad;
	return BG_OK;
}

extern "C" int bg_search_bunches_into(bg_ctx *c, const bg_reads *R, uint32_t qbunch, int heuristic, int skip_bad, int mode,
		uint16_t *best_inout, bg_xhit *hits, uint64_t cap, uint64_t *nhits) {
	if (!c || !R || !nhits || (!hits && cap)) return fail(BG_EINVAL, "bg_search_bunches_into: null argument");
	if (!c->acx_n) return fail(BG_EINVAL, "bg_search_bunches_into: no accelerator loaded (bg_load_acx)");
	if (!R->reads || !R->len || !R->budget || !R->strand || !R->nreads || !R->nq) return fail(BG_EINVAL, "bg_search_bunches_into: null or empty read arrays");
	if (R->flags != BG_R_PACKED2) return fail(BG_EINVAL, "bg_search_bunches_into: reads must be BG_R_PACKED2 (plain bases)");
	if (!qbunch || qbunch > BG_RUN_MAX) return fail(BG_EINVAL, "bg_search_bunches_into: bunch size %u (must be 1..%d)", qbunch, BG_RUN_MAX);
	CU(cudaSetDevice(c->device));
	c->kind = WORK_NONE; c->ran = false;
	const uint32_t nq = R->nq, nr = R->nreads, nbunch = (nq + qbunch - 1) / qbunch;
	cudaStream_t st = c->stream;
	uint64_t total = 0; uint32_t maxlen = 0;
	for (uint32_t r = 0; r < nr; ++r) { total += R->len[r]; maxlen = std::max<uint32_t>(maxlen, R->len[r]); }
	const uint64_t rbytes = (total + 3) / 4, ncodes_max = (uint64_t)nq * (((uint64_t)maxlen + 15) & ~15ull);
	if (c->d_rlen.need(nr) || c->d_rbud.need(nr) || c->d_strand.need(nq) || c->d_rl64.need((size_t)nr + 1) || c->d_sl64.need((size_t)nq + 1) || c->d_roff.need((size_t)nr + 1) ||
	    c->d_qoff.need((size_t)nq + 1) || c->d_qi.need(nq) || c->d_best.need(nr) || c->d_best16.need(nr) || c->d_counters.need(64) || c->d_cells.need(8) ||
	    c->d_packed.need(rbytes + 32) || c->d_codes.need(ncodes_max + 32) || c->d_qnib.need(ncodes_max / 8 + 3ull * nq + 8)) return BG_ENOMEM;
	size_t tmp1 = 0, tmp2 = 0;
	CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp1, c->d_rl64.p, c->d_roff.p, (int)nr + 1, st));
	CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, c->d_sl64.p, (unsigned long long *)c->d_qoff.p, (int)nq + 1, st));
	if (c->d_sort_tmp.need(std::max(tmp1, tmp2) + 16)) return BG_ENOMEM;
	// ---- strands on the device (as bg_align_bunches_into) ----
	CU(cudaMemcpyAsync(c->d_rlen.p, R->len, (size_t)nr * 2, cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(c->d_rbud.p, R->budget, (size_t)nr * 2, cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(c->d_strand.p, R->strand, (size_t)nq * 4, cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(c->d_packed.p, R->reads, rbytes, cudaMemcpyHostToDevice, st));
	CU(cudaMemsetAsync(c->d_counters.p, 0, 256, st));
	k_compact_rlen<<<(nr + 256) / 256, 256, 0, st>>>(c->d_rlen.p, nr, c->d_rl64.p);
	k_compact_slen<<<(nq + 256) / 256, 256, 0, st>>>(c->d_rlen.p, nr, c->d_strand.p, nq, c->d_sl64.p, c->d_counters.p);
	CU(cub::DeviceScan::ExclusiveSum(c->d_sort_tmp.p, tmp1, c->d_rl64.p, c->d_roff.p, (int)nr + 1, st));
	CU(cub::DeviceScan::ExclusiveSum(c->d_sort_tmp.p, tmp2, c->d_sl64.p, (unsigned long long *)c->d_qoff.p, (int)nq + 1, st));
	const bool prep2 = R->flags == BG_R_PACKED2 && !getenv("BURST_B200_NO_PREP2");
	if (!prep2) k_compact_codes<<<(unsigned)(((uint64_t)nq * 8 + 255) / 256), 256, 0, st>>>(c->d_packed.p, R->flags, c->d_roff.p, c->d_rlen.p, c->d_strand.p, (const unsigned long long *)c->d_qoff.p, nq, nr, c->d_codes.p);
	k_compact_qinfo<<<(nq + 255) / 256, 256, 0, st>>>((const unsigned long long *)c->d_qoff.p, c->d_rlen.p, c->d_rbud.p, c->d_strand.p, nq, nr, c->d_qi.p, c->d_counters.p + 16, c->d_counters.p);
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(c->h_pinned, c->d_counters.p, 256, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	if (c->h_pinned[C_ERR]) return fail(BG_EINVAL, "bg_search_bunches_into: malformed batch (strands must name reads < nreads, read lengths >= 1, budgets <= 254)");
	c->SL = choose_layout(c, c->h_pinned + 16, nq);
	c->mstage = stage_len(maxlen); c->wide_possible = long_queries(maxlen); c->len_hint = maxlen;
	if (prep2) k_compact_prep2<<<(unsigned)std::min<uint64_t>(((uint64_t)nq * 8 + 255) / 256, (uint64_t)c->sms * 8), 256, 0, st>>>((const uint32_t *)c->d_packed.p, c->d_roff.p, c->d_rlen.p, c->d_rbud.p, c->d_strand.p, (const unsigned long long *)c->d_qoff.p, nq, nr,
		c->d_codes.p, c->d_qi.p, c->SL, c->d_qnib.p, c->d_counters.p + 9, nullptr);
	else k_qprep<<<(nq + 127) / 128, 128, 0, st>>>(c->d_codes.p, c->d_qi.p, nq, c->SL, c->d_qnib.p, c->d_counters.p + 9, nullptr);
	CU(cudaGetLastError());
	c->nq = nq; c->nslots = nr;
	// ---- candidates -> runs on the device ----
	CandArgs G;
	G.qi = c->d_qi.p; G.qnib = c->d_qnib.p; G.nq = nq; G.qbunch = qbunch; G.nbunch = nbunch;
	G.acx_off = c->d_acx_off.p; G.post = c->d_post.p; G.big = c->acx_big; G.N = c->acx_n; G.num_clumps = c->acx_clumps;
	G.bad = c->d_bad.p; G.nbad = c->acx_nbad; G.heur = heuristic; G.skip_bad = skip_bad;
	G.cnt = c->d_cg_cnt.p; G.first = c->d_cg_first.p; G.cache = c->d_cg_cache.p; G.counters = c->d_counters.p;
	{ const uint32_t words = maxlen >= (uint32_t)c->acx_n ? maxlen - c->acx_n + 1 : 1; uint32_t pm = 1024; while (pm < qbunch * words) pm <<= 1;
	  if (pm > 8192) return fail(BG_EINVAL, "bg_search_bunches_into: reads of %u bases need %u (word, query) pairs per bunch, the device path holds 8192", maxlen, qbunch * words);
	  G.pmax = pm; G.cmax = std::min<uint32_t>(pm, 4096); }
	const size_t cg_smem = (size_t)G.pmax * 8 + (size_t)G.cmax * 8;
	if (cg_smem > 48 * 1024) CU(cudaFuncSetAttribute(k_candgen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cg_smem));
	if (!c->runs_cap) c->runs_cap = 1u << 20;
	c->runs_cap = (uint32_t)std::min<uint64_t>((1ull << 28) - 2, std::max<uint64_t>(c->runs_cap, (uint64_t)nbunch * (12 + (skip_bad ? 0 : c->acx_nbad))));
	uint64_t nruns = 0;
	for (int attempt = 0; attempt < 3; ++attempt) {
		if (c->d_runs.need((size_t)c->runs_cap + 1)) return BG_ENOMEM;
		G.runs = c->d_runs.p; G.runs_cap = c->runs_cap;
		CU(cudaMemsetAsync(c->d_counters.p + C_RUNS, 0, 4, st));
		k_candgen<<<std::min<uint32_t>(nbunch, c->cg_blocks), 128, cg_smem, st>>>(G);
		CU(cudaGetLastError());
		CU(cudaMemcpyAsync(c->h_pinned, c->d_counters.p, 64, cudaMemcpyDeviceToHost, st));
		CU(cudaStreamSynchronize(st));
		if (c->h_pinned[C_ERR]) return fail(BG_EOVERFLOW, "bg_search_bunches_into: a bunch has more candidates or words than the device tables hold (flag %#x); use the host candidate lists for this batch", c->h_pinned[C_ERR]);
		nruns = c->h_pinned[C_RUNS];
		if (nruns <= c->runs_cap) break;
		if (nruns >= (1ull << 28)) return fail(BG_EINVAL, "bg_search_bunches_into: %llu runs in one batch (limit 2^28-1); use smaller batches", (unsigned long long)nruns);
		c->runs_cap = (uint32_t)(nruns + nruns / 8);
	}
	if (nruns > c->runs_cap) return fail(BG_EOVERFLOW, "bg_search_bunches_into: run list kept overflowing");
	c->kind = WORK_RUNS; c->nruns = nruns; c->ntasks = nruns * BG_RUN_MAX; c->ntiles = 0;
	CU(cudaMemsetAsync(c->d_counters.p, 0, 32, st));                   // (k_qprep's statistics at [9..11] stay)
	int rc = finish_upload(c); if (rc) return rc;
	rc = bg_batch_run(c, mode, best_inout); if (rc) return rc;
	uint64_t n = 0;
	rc = bg_batch_count(c, &n); if (rc) return rc;
	*nhits = n;
	if (n > cap) return fail(BG_EOVERFLOW, "bg_search_bunches_into: %llu hits, room for %llu", (unsigned long long)n, (unsigned long long)cap);
	rc = sort_hits(c, (uint32_t)n); if (rc) return rc;
	if (n) {
		if (c->d_xhits.need(n)) return BG_ENOMEM;
		k_xhits<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->d_hits_sorted.p, c->d_runs.p, (uint32_t)n, c->d_xhits.p);
		CU(cudaGetLastError());
		CU(cudaMemcpyAsync(hits, c->d_xhits.p, n * sizeof(bg_xhit), cudaMemcpyDeviceToHost, st));
	}
	if (best_inout) {
		k_best16<<<(nr + 255) / 256, 256, 0, st>>>(c->d_best.p, c->d_best16.p, nr);
		CU(cudaMemcpyAsync(best_inout, c->d_best16.p, (size_t)nr * 2, cudaMemcpyDeviceToHost, st));
	}
	CU(cudaStreamSynchronize(st));
	return BG_OK;
}

static int finish_align(bg_ctx *c, int mode, uint16_t *best_inout, bg_hit **hits, uint64_t *nhits) {
	int rc = bg_batch_run(c, mode, best_inout); if (rc) return rc;
	uint64_t n = 0;
	rc = bg_batch_count(c, &n); if (rc) return rc;
	bg_hit *h = (bg_hit *)malloc((n ? n : 1) * sizeof(bg_hit));
	if (!h) return fail(BG_ENOMEM, "malloc hits");
	rc = bg_batch_download(c, h, n, best_inout);
	if (rc) { free(h); return rc; }
	*hits = h; *nhits = n;
	return BG_OK;
}

extern "C" int bg_align_batch(bg_ctx *c, const bg_queries *Q, const bg_task *tasks, uint64_t ntasks, int mode,
		uint16_t *best_inout, bg_hit **hits, uint64_t *nhits) {
	if (!hits || !nhits) return fail(BG_EINVAL, "bg_align_batch: null output");
	int rc = bg_batch_upload(c, Q, tasks, ntasks); if (rc) return rc;
	return finish_align(c, mode, best_inout, hits, nhits);
}
extern "C" int bg_align_runs(bg_ctx *c, const bg_queries *Q, const bg_run *runs, uint64_t nruns, int mode,
		uint16_t *best_inout, bg_hit **hits, uint64_t *nhits) {
	if (!hits || !nhits) return fail(BG_EINVAL, "bg_align_runs: null output");
	if (pipeline_applies(c, Q, nruns) && runs) {
		// hits are bounded by the survivors: size the host buffer after a first pass would cost a second one, so take the
		// one-call path with a generous buffer and shrink it
		uint64_t cap = std::max<uint64_t>(1024, (uint64_t)Q->nq * 2), n = 0;
		for (int attempt = 0; attempt < 2; ++attempt) {
			bg_hit *h = (bg_hit *)malloc(cap * sizeof(bg_hit));
			if (!h) return fail(BG_ENOMEM, "malloc hits");
			int rc = bg_align_runs_into(c, Q, runs, nruns, mode, best_inout, h, cap, &n);
			if (rc == BG_OK) { *hits = h; *nhits = n; return BG_OK; }
			free(h);
			if (rc != BG_EOVERFLOW || n <= cap) return rc;
			cap = n;
		}
		return fail(BG_EOVERFLOW, "bg_align_runs: hit buffer kept overflowing");
	}
	int rc = bg_batch_upload_runs(c, Q, runs, nruns); if (rc) return rc;
	return finish_align(c, mode, best_inout, hits, nhits);
}
extern "C" void bg_free_hits(bg_hit *h) { free(h); }
