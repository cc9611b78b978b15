// burst_b200.cu -- sm_100a kernels + C ABI (include/burst_b200.h) of the B200 alignment engine.
//
// What the reference does per (query, clump) pair:
//   pass 1  aded_mat16L / aded_mat16   burst.c:1003-1204   16-lane banded edit distance
//   pass 2  reScoreM_mat16             burst.c:713-886     (score, shift, shiftR) + end column
// and how this file restructures it for the GPU (DESIGN.md has the full argument):
//   k_query_prep   per query, 16 match bit-vectors over its first P <= 32 rows
//   k_filter       one thread per (task, lane): bit-parallel (Myers/Hyyro) semi-global DP of
//                  the query's first P rows over every column of the lane.  Any alignment
//                  with <= k errors has a prefix with <= k errors, so columns whose row-P value
//                  is <= k ("seeds") cover every cell of every <= k alignment within +-k
//                  diagonals.  >92% of the reference's pass-1 calls die in these rows.
//   k_extend       one thread per surviving (task, lane): exact banded DP over the hull of the
//                  seed diagonals, carrying the reference's pass-2 triple packed in one 32-bit
//                  key so that diag/up/left selection with its tie-break order is a 3-operand
//                  min/add (DPX VIADDMNMX / VIMNMX3).  Yields pass 1's distance and pass 2's
//                  (numGapQ, numGapR, finalPos) in one sweep; atomicMin keeps the per-slot best.
//   k_select       keeps the lanes the reference would have kept (burst.c:4219-4229).
// Values <= budget are exact and identical to the reference's saturating u8 arithmetic because
// every cell > maxED is treated as absent there too (burst.c:1053-1054, 802-803).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <algorithm>
#include <vector>
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
struct QInfo {            // one query of the batch
	uint64_t off;         // into codes
	uint32_t len;
	uint32_t slot;
	uint16_t k;           // budget (Emac)
	uint16_t P;           // rows covered by the prefix filter = min(32, len)
};
struct Surv {             // one (task, lane) that survived the prefix filter
	uint32_t task;
	int32_t  lo;          // lowest diagonal (x - y) of the band
	uint32_t w_lane;      // band width << 8 | lane
	uint32_t scratch;     // offset into the global band scratch (generic kernel only)
};
struct Res { uint32_t a, b; };   // a = ed | gap_q << 8 | gap_r << 16 | valid << 31 ; b = final_pos

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
		uint4 *__restrict__ out, uint32_t first, uint64_t in_base) {
	uint32_t c = first + blockIdx.x;
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
	}
}

// ---------------------------------------------------------------------------------------------
// Query prep: Peq[q][c] bit (32-P+y-1) = 1 iff row y (1-based) of the query matches reference
// code c (S == 0).  The pattern is left-aligned so that row P sits in bit 31; the unused low
// 32-P bits are set for every code: with Pv = Mv = 0 there they behave as extra copies of the
// all-zero row 0 of the semi-global matrix and never generate a carry.
// ---------------------------------------------------------------------------------------------
__global__ void k_query_prep(const uint8_t *__restrict__ codes, const QInfo *__restrict__ qi,
		const uint32_t *__restrict__ Sterm, uint32_t nq, uint32_t *__restrict__ peq) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t q = i >> 4, c = i & 15;
	if (q >= nq) return;
	QInfo Q = qi[q];
	uint32_t P = Q.P, m = P < 32 ? (1u << (32 - P)) - 1 : 0;
	const uint8_t *s = codes + Q.off;
	for (uint32_t y = 0; y < P; ++y)
		if (Sterm[s[y] * 16 + c] == 0) m |= 1u << (32 - P + y);
	peq[i] = m;
}

// ---------------------------------------------------------------------------------------------
// Phase A: bit-parallel prefix filter.  128 threads = 8 tasks x 16 lanes.
// ---------------------------------------------------------------------------------------------
struct FilterArgs {
	const uint4 *db; const uint64_t *clump_off; const uint32_t *clump_len;
	const QInfo *qi; const uint32_t *peq; const bg_task *tasks;
	uint64_t ntasks; uint32_t nq, first_clump, num_clumps;
	Surv *surv; uint32_t surv_cap; uint32_t *counters;   // [0] survivors, [1] scratch words, [2] hits
	uint32_t two;        // the constant 2, passed at run time so ptxas keeps IMAD.HI / IMAD.WIDE (FMA pipe)
	uint32_t c16;        // the constant 16, same reason
};

__device__ __forceinline__ void task_of(const FilterArgs &A, uint64_t t, uint32_t &q, uint32_t &c) {
	if (A.tasks) { bg_task T = A.tasks[t]; q = T.query; c = T.clump; }
	else { q = (uint32_t)(t % A.nq); c = (uint32_t)(t / A.nq) + A.first_clump; }
}

// Hyyro's formulation of Myers' bit-vector step; the text character is one reference base.
// Integer-pipe budget per column (ncu: the kernel is bound by the ALU pipe, LOP3/SHF/ISETP issue
// at half rate): the seven 3-input logic ops below are irreducible, so everything that can run
// on the FMA pipe instead is written as a multiply-add: the two shifts are x+x, the nibble
// extraction is a mul.hi, and the row-P score is kept as two mad.hi accumulators (bit 31 of
// Ph / Mh) that are only compared once per 8 columns.
__device__ __forceinline__ uint32_t madhi(uint32_t a, uint32_t b, uint32_t c) {
	uint32_t d; asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ uint32_t mulhi(uint32_t a, uint32_t b) {
	uint32_t d; asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
#define MYERS_STEP(Eq)                                          \
	const uint32_t Xv = (Eq) | Mv;                              \
	const uint32_t Xh = ((((Eq) & Pv) + Pv) ^ Pv) | (Eq);       \
	uint32_t Ph = Mv | ~(Xh | Pv);                              \
	uint32_t Mh = Pv & Xh;

template <int V>
__global__ void __launch_bounds__(128) k_filter(FilterArgs A) {
	__shared__ uint32_t sPeq[8][16];
	const uint32_t slot = threadIdx.x >> 4, lane = threadIdx.x & 15;
	const uint64_t t = (uint64_t)blockIdx.x * 8 + slot;
	bool valid = t < A.ntasks;
	uint32_t q = 0, c = 0;
	if (valid) {
		task_of(A, t, q, c);
		c -= A.first_clump;
		valid = c < A.num_clumps;              // other shards' clumps are skipped
	}
	sPeq[slot][lane] = valid ? A.peq[(size_t)q * 16 + lane] : 0;
	__syncwarp();
	if (!valid) return;
	const QInfo Q = A.qi[q];
	const int P = Q.P, k = Q.k;
	const uint32_t L = A.clump_len[c];
	const uint4 *base = A.db + A.clump_off[c] + lane;
	const char *eq = (const char *)sPeq[slot];

	uint32_t Pv = P < 32 ? ~0u << (32 - P) : ~0u, Mv = 0;
	uint32_t cP = (uint32_t)P, cM = 0;          // row-P value = cP - cM
	int lo = INT32_MAX, hi = INT32_MIN;
	const uint32_t nwords = (L + 7) >> 3;
	uint4 w4 = base[0];
	for (uint32_t wi0 = 0; wi0 < nwords; wi0 += 4) {
		uint4 nx = w4;
		if (wi0 + 4 < nwords) nx = base[(size_t)((wi0 >> 2) + 1) * 16];
		const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
		#pragma unroll
		for (int wi = 0; wi < 4; ++wi) {
			if (wi0 + wi >= nwords) break;
			const uint32_t w = ws[wi];
			const uint32_t Pv0 = Pv, Mv0 = Mv;
			const int s0 = (int)(cP - cM);
			if (V == 5) {
				// as V == 4 but the shared address is formed on the ALU pipe (shift, and-or): no IMAD.HI
				const uint32_t eqs = (uint32_t)__cvta_generic_to_shared(eq);
				#pragma unroll
				for (int j = 0; j < 8; ++j) {
					const uint32_t sh = j == 0 ? w << 2 : w >> (4 * j - 2);
					uint32_t Eq;
					asm volatile("ld.shared.u32 %0, [%1];" : "=r"(Eq) : "r"((sh & 0x3Cu) | eqs));
					MYERS_STEP(Eq)
					Ph += Ph; Mh += Mh;
					Pv = Mh | ~(Xv | Ph);
					Mv = Ph & Xv;
				}
				cP = (uint32_t)__popc(Pv); cM = (uint32_t)__popc(Mv);
			} else if (V == 4) {
				// ALU pipe: only the seven logic ops.  FMA pipe: nibble -> shared address (shl, mul.hi, mad),
				// the add and the two shifts.  The row-P value is not tracked at all: it is
				// popc(Pv) - popc(Mv) (sum of the vertical deltas over the pattern rows), read once per word.
				const uint32_t eqs = (uint32_t)__cvta_generic_to_shared(eq);
				#pragma unroll
				for (int j = 0; j < 8; ++j) {
					const uint32_t code = mulhi(j == 7 ? w : w << (28 - 4 * j), A.c16);
					uint32_t Eq;
					asm volatile("ld.shared.u32 %0, [%1];" : "=r"(Eq) : "r"(code * 4u + eqs));
					MYERS_STEP(Eq)
					Ph += Ph; Mh += Mh;
					Pv = Mh | ~(Xv | Ph);
					Mv = Ph & Xv;
				}
				cP = (uint32_t)__popc(Pv); cM = (uint32_t)__popc(Mv);
			} else
			#pragma unroll
			for (int j = 0; j < 8; ++j) {
				const uint32_t off = (j == 0 ? w * 4u : mulhi(w, 1u << (34 - 4 * j))) & 0x3Cu;   // code * 4
				const uint32_t Eq = *(const uint32_t *)(eq + off);
				MYERS_STEP(Eq)
				if (V == 1) { cP = madhi(Ph, 2u, cP); cM = madhi(Mh, 2u, cM); Ph += Ph; Mh += Mh; }      // ptxas: LEA.HI (ALU)
				else if (V == 2) { cP = madhi(Ph, A.two, cP); cM = madhi(Mh, A.two, cM); Ph += Ph; Mh += Mh; }  // IMAD.HI (FMA)
				else {                                               // IMAD.WIDE: shift and bit 31 in one
					uint64_t p2, m2;
					asm("mul.wide.u32 %0, %1, %2;" : "=l"(p2) : "r"(Ph), "r"(A.two));
					asm("mul.wide.u32 %0, %1, %2;" : "=l"(m2) : "r"(Mh), "r"(A.two));
					Ph = (uint32_t)p2; Mh = (uint32_t)m2; cP += (uint32_t)(p2 >> 32); cM += (uint32_t)(m2 >> 32);
				}
				Pv = Mh | ~(Xv | Ph);                                // row 0 is all zero: no carry-in
				Mv = Ph & Xv;
			}
			// The row-P value moves by at most 1 per column, so inside these 8 columns it cannot
			// drop below (s0 + s1 - 8) / 2.  Only then is a seed (value <= k) possible: redo the
			// word column by column from the saved state.
			const int s1 = (int)(cP - cM);
			if (s0 + s1 - 8 <= 2 * k) {
				uint32_t pv = Pv0, mv = Mv0; int score = s0;
				#pragma unroll 1
				for (int j = 0; j < 8; ++j) {
					const uint32_t Eq = *(const uint32_t *)(eq + (((w >> (4 * j)) & 15u) << 2));
					const uint32_t xv = Eq | mv;
					const uint32_t xh = (((Eq & pv) + pv) ^ pv) | Eq;
					uint32_t ph = mv | ~(xh | pv), mh = pv & xh;
					score += (int)(ph >> 31) - (int)(mh >> 31);
					ph <<= 1; mh <<= 1;
					pv = mh | ~(xv | ph); mv = ph & xv;
					const int x = (int)((wi0 + wi) * 8 + j) + 1;
					if (score <= k && x <= (int)L) {                 // seed: D[P][x] <= k
						const int d = x - P;
						lo = min(lo, d - k); hi = max(hi, d + k);
					}
				}
			}
		}
		w4 = nx;
	}
	if (lo <= hi) {
		const uint32_t W = (uint32_t)(hi - lo + 1);
		uint32_t scratch = 0;
		if (W > 64) scratch = atomicAdd(&A.counters[1], W);
		const uint32_t i = atomicAdd(&A.counters[0], 1u);
		if (i < A.surv_cap) {
			Surv s; s.task = (uint32_t)t; s.lo = lo; s.w_lane = (W << 8) | lane; s.scratch = scratch;
			A.surv[i] = s;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Phase B: exact banded DP with the pass-2 triple.
// ---------------------------------------------------------------------------------------------
struct ExtendArgs {
	const uint32_t *dbw; const uint64_t *clump_off; const uint32_t *clump_len;
	const uint8_t *codes; const QInfo *qi; const bg_task *tasks;
	uint32_t nq, first_clump;
	const Surv *surv; uint32_t surv_cap; const uint32_t *counters;
	Res *res; uint32_t *best; const uint32_t *Sterm;   // Sterm[q*16+r] = S << 22
	uint32_t *scratch; uint32_t scratch_cap;
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

template <int WMAX>
__global__ void __launch_bounds__(128) k_extend(ExtendArgs A) {
	__shared__ uint32_t sS[256];
	for (int i = threadIdx.x; i < 256; i += blockDim.x) sS[i] = A.Sterm[i];
	__syncthreads();
	const uint32_t nsurv = min(A.counters[0], A.surv_cap);
	unsigned long long cells = 0;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nsurv; i += gridDim.x * blockDim.x) {
		const Surv sv = A.surv[i];
		const uint32_t W = sv.w_lane >> 8, lane = sv.w_lane & 255;
		// class dispatch: this instantiation takes bands that fit WMAX but not WMAX/2
		if (WMAX == 0 ? (W <= 64) : (W > (uint32_t)WMAX || (WMAX > 8 && W <= (uint32_t)WMAX / 2))) continue;
		uint32_t q, c;
		if (A.tasks) { bg_task T = A.tasks[sv.task]; q = T.query; c = T.clump; }
		else { q = sv.task % A.nq; c = sv.task / A.nq + A.first_clump; }
		c -= A.first_clump;
		const QInfo Q = A.qi[q];
		const uint32_t m = Q.len, L = A.clump_len[c];
		const uint8_t *qs = A.codes + Q.off;
		const uint32_t *lanew = A.dbw + A.clump_off[c] * 4 + lane * 4;
		uint32_t k = Q.k;
		if (A.mode == BG_MODE_MIN) k = min(k, A.best[Q.slot]);
		uint32_t inf = (k + 1) << 22;
		const int lo = sv.lo;
		constexpr int WB = WMAX ? WMAX : 1;
		const int Wd = WMAX ? WMAX : (int)W;                 // cells per row actually swept
		uint32_t a[WB];                                      // band, register resident when WMAX > 0
		uint32_t win[WMAX ? (WMAX + 7) / 8 : 1];             // codes of columns x0 .. x0+WMAX-1, one nibble each
		uint32_t *g = A.scratch + sv.scratch;                // generic path: band in global scratch
		if (WMAX == 0 && (uint64_t)sv.scratch + W > A.scratch_cap) { A.res[i].a = 0; continue; }

		// row 0: zero for columns 0..L (burst.c:4052 calloc / 723-725), absent elsewhere
		if (WMAX) {
			#pragma unroll
			for (int d = 0; d < WB; ++d) { const int x = lo + d; a[d] = (x >= 0 && x <= (int)L) ? KEY_ZERO : inf; }
			#pragma unroll
			for (int j = 0; j < (WB + 7) / 8; ++j) win[j] = 0;
			#pragma unroll
			for (int d = 0; d < WB; ++d) win[d >> 3] |= fetch_code(lanew, (uint32_t)(lo + d - 1), L) << (4 * (d & 7));
		} else {
			for (int d = 0; d < Wd; ++d) { const int x = lo + d; g[d] = (x >= 0 && x <= (int)L) ? KEY_ZERO : inf; }
		}

		bool dead = false;
		uint32_t y = 1;
		for (; y <= m; ++y) {
			const int x0 = (int)y + lo;                      // column (1-based) of band cell 0 in row y
			const uint32_t *Srow = sS + qs[y - 1] * 16;
			uint32_t rowmin = KEY_NONE, left = inf;
			if (WMAX) {
				// slide the code window by one column
				const uint32_t nc = fetch_code(lanew, (uint32_t)(x0 + WB - 2), L);
				#pragma unroll
				for (int j = 0; j < (WB + 7) / 8 - 1; ++j) win[j] = __funnelshift_r(win[j], win[j + 1], 4);
				win[(WB + 7) / 8 - 1] = (win[(WB + 7) / 8 - 1] >> 4) | (nc << (4 * ((WB - 1) & 7)));
				if (x0 >= 1 && x0 + WB - 1 <= (int)L) {      // interior row: every cell and predecessor is inside the matrix
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
			} else {
				uint32_t diag = g[0];
				for (int d = 0; d < Wd; ++d) {
					const int x = x0 + d;
					const uint32_t up = d + 1 < Wd ? g[d + 1] : inf;
					const uint32_t st = Srow[fetch_code(lanew, (uint32_t)(x - 1), L)];
					uint32_t v = cell(diag, up, left, st, inf);
					if (x < 0 || x > (int)L) v = inf;
					else if (x == 0) v = y <= k ? key_col0(y) : inf;
					diag = up; g[d] = v; left = v; rowmin = min(rowmin, v);
				}
			}
			if (rowmin >= inf) { dead = true; break; }       // every lane cell > maxED: the reference truncates (burst.c:1062-1065)
			if ((y & 15) == 0 && A.mode == BG_MODE_MIN) {    // tighten Emac as better hits land (burst.c:4159, 4220)
				k = min(k, A.best[Q.slot]); inf = (k + 1) << 22;
			}
		}
		cells += (unsigned long long)(dead ? y : m) * (unsigned)Wd;
		uint32_t out = 0, fp = 0;
		if (!dead) {
			// last-row selection, left to right (burst.c:826-842, 863-883)
			uint32_t bk = KEY_NONE >> 11, bshr = 0;
			auto scan = [&](int d, uint32_t v) {
				const int x = (int)m + lo + d;
				if (x < 1 || x > (int)L) return;
				const uint32_t kk = v >> 11;
				if (kk < bk) { bk = kk; bshr = v & 0x1FF; fp = (uint32_t)x; }
				else if (kk == bk) fp = (uint32_t)x;
			};
			if (WMAX) {
				#pragma unroll
				for (int d = 0; d < WB; ++d) scan(d, a[d]);
			} else for (int d = 0; d < Wd; ++d) scan(d, g[d]);
			const uint32_t ed = bk >> 11, sh = 2047u - (bk & 2047u);
			if (bk != (KEY_NONE >> 11) && ed <= k) {
				out = ed | (sh << 8) | (bshr << 16) | (1u << 31);
				atomicMin(&A.best[Q.slot], ed);
			}
		}
		Res r; r.a = out; r.b = fp;
		A.res[i] = r;
	}
	if (cells) atomicAdd(A.band_cells, cells);
}

// ---------------------------------------------------------------------------------------------
// Phase C: keep what the reference keeps.
// ---------------------------------------------------------------------------------------------
__global__ void k_select(const Surv *__restrict__ surv, const Res *__restrict__ res, const bg_task *__restrict__ tasks,
		const QInfo *__restrict__ qi, uint32_t nq, const uint32_t *__restrict__ best, uint32_t *counters,
		uint32_t surv_cap, bg_hit *__restrict__ hits, int mode) {
	const uint32_t nsurv = min(counters[0], surv_cap);
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nsurv; i += gridDim.x * blockDim.x) {
		const Res r = res[i];
		if (!(r.a >> 31)) continue;
		const Surv sv = surv[i];
		const uint32_t q = tasks ? tasks[sv.task].query : sv.task % nq;
		const uint32_t ed = r.a & 255;
		if (mode == BG_MODE_MIN && ed != best[qi[q].slot]) continue;     // burst.c:4229, 4497
		const uint32_t j = atomicAdd(&counters[2], 1u);
		bg_hit h; h.task = sv.task; h.lane = (uint8_t)(sv.w_lane & 255); h.ed = (uint8_t)ed;
		h.gap_q = (uint8_t)(r.a >> 8); h.gap_r = (uint8_t)(r.a >> 16); h.final_pos = r.b;
		hits[j] = h;
	}
}

__global__ void k_init_best(uint32_t *best, const uint16_t *in, uint32_t n) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) best[i] = in ? in[i] : 0xFFFFu;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename T> struct DBuf {
	T *p = nullptr; size_t cap = 0;
	int need(size_t n) {
		if (n <= cap) return 0;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = n + n / 8 + 64;
		cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
		if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(BG_ENOMEM, "cudaMalloc(%zu bytes): %s", want * sizeof(T), cudaGetErrorString(e)); }
		cap = want; return 0;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct bg_ctx {
	int device = 0;
	cudaStream_t stream = nullptr; bool own_stream = false;
	int sms = 148;
	// scoring
	uint8_t S[256];
	DBuf<uint32_t> d_sterm;
	// DB
	DBuf<uint4> d_db; DBuf<uint64_t> d_clump_off; DBuf<uint32_t> d_clump_len;
	std::vector<uint32_t> clump_len;
	uint32_t num_clumps = 0, first_clump = 0;
	// batch
	DBuf<uint8_t> d_codes; DBuf<QInfo> d_qi; DBuf<uint32_t> d_peq; DBuf<bg_task> d_tasks;
	DBuf<uint32_t> d_best; DBuf<uint16_t> d_best16;
	DBuf<Surv> d_surv; DBuf<Res> d_res; DBuf<bg_hit> d_hits; DBuf<uint32_t> d_scratch;
	DBuf<uint32_t> d_counters; DBuf<unsigned long long> d_cells;
	std::vector<QInfo> h_qi;
	uint32_t nq = 0, nslots = 0; uint64_t ntasks = 0; bool have_tasks = false;
	uint32_t surv_cap = 0;
	int last_mode = 0; std::vector<uint16_t> last_best_in; bool have_best_in = false;
	bg_stats stats;
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
	uint32_t h_counters[4] = {0, 0, 0, 0};
	bool ran = false;
	bool by_runs = false; std::vector<uint32_t> run_key;      // run-list batches: task index -> run * BG_RUN_MAX + i
};

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
	memset(&c->stats, 0, sizeof(c->stats));
	bg_default_scoring(1, c->S);
	*out = c;
	int rc = bg_set_scoring(c, c->S);
	if (rc) { return rc; }
	return BG_OK;
}

extern "C" void bg_free(bg_ctx *c) {
	if (!c) return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	c->d_sterm.release(); c->d_db.release(); c->d_clump_off.release(); c->d_clump_len.release();
	c->d_codes.release(); c->d_qi.release(); c->d_peq.release(); c->d_tasks.release();
	c->d_best.release(); c->d_best16.release(); c->d_surv.release(); c->d_res.release();
	c->d_hits.release(); c->d_scratch.release(); c->d_counters.release(); c->d_cells.release();
	for (int i = 0; i < 4; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
	if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

extern "C" int bg_set_stream(bg_ctx *c, void *s) {
	if (!c) return fail(BG_EINVAL, "null ctx");
	if (c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
	c->stream = (cudaStream_t)s; c->own_stream = false;
	return BG_OK;
}

extern "C" int bg_set_scoring(bg_ctx *c, const uint8_t S[256]) {
	if (!c || !S) return fail(BG_EINVAL, "bg_set_scoring: null argument");
	CU(cudaSetDevice(c->device));
	memcpy(c->S, S, 256);
	uint32_t st[256];
	for (int i = 0; i < 256; ++i) st[i] = (uint32_t)S[i] << 22;
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
	c->clump_len.assign(clump_len, clump_len + num_clumps);
	c->num_clumps = num_clumps; c->first_clump = first_clump;
	if (c->d_db.need(out_off[num_clumps])) return BG_ENOMEM;
	if (c->d_clump_off.need(num_clumps + 1) || c->d_clump_len.need(num_clumps)) return BG_ENOMEM;
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
		k_relayout<<<j - i, 128, 0, c->stream>>>(stage.p, d_in_off.p, c->d_clump_off.p, c->d_clump_len.p, c->d_db.p, i, in_off[i]);
		CU(cudaGetLastError());
		CU(cudaStreamSynchronize(c->stream));
		i = j;
	}
	stage.release(); d_in_off.release();
	return BG_OK;
}

extern "C" int bg_batch_upload(bg_ctx *c, const bg_queries *Q, const bg_task *tasks, uint64_t ntasks) {
	if (!c || !Q) return fail(BG_EINVAL, "bg_batch_upload: null argument");
	if (!c->num_clumps) return fail(BG_EINVAL, "bg_batch_upload: no database loaded");
	if (!Q->nq) return fail(BG_EINVAL, "bg_batch_upload: empty query batch");
	CU(cudaSetDevice(c->device));
	c->h_qi.resize(Q->nq);
	for (uint32_t i = 0; i < Q->nq; ++i) {
		uint64_t len = Q->offset[i + 1] - Q->offset[i];
		if (!len || len > 0x7FFFFFFF) return fail(BG_EINVAL, "bg_batch_upload: query %u has length %llu", i, (unsigned long long)len);
		if (Q->budget[i] > 254) return fail(BG_EINVAL, "bg_batch_upload: budget %u of query %u exceeds 254 (burst.c:3076)", Q->budget[i], i);
		if (Q->slot[i] >= Q->nslots) return fail(BG_EINVAL, "bg_batch_upload: slot %u of query %u out of range", Q->slot[i], i);
		QInfo &q = c->h_qi[i];
		q.off = Q->offset[i]; q.len = (uint32_t)len; q.slot = Q->slot[i]; q.k = Q->budget[i];
		q.P = (uint16_t)std::min<uint64_t>(32, len);
	}
	uint64_t ncodes = Q->offset[Q->nq];
	if (!tasks) {
		ntasks = (uint64_t)Q->nq * c->num_clumps;
	} else {
		for (uint64_t t = 0; t < ntasks; ++t)
			if (tasks[t].query >= Q->nq) return fail(BG_EINVAL, "bg_batch_upload: task %llu names query %u of %u", (unsigned long long)t, tasks[t].query, Q->nq);
	}
	if (ntasks >= (1ull << 32)) return fail(BG_EINVAL, "bg_batch_upload: %llu tasks in one batch (limit 2^32-1); split the query batch", (unsigned long long)ntasks);
	if (c->d_codes.need(ncodes + 16) || c->d_qi.need(Q->nq) || c->d_peq.need((size_t)Q->nq * 16) ||
	    c->d_best.need(Q->nslots) || c->d_best16.need(Q->nslots) || c->d_counters.need(4) || c->d_cells.need(1)) return BG_ENOMEM;
	if (tasks && c->d_tasks.need(ntasks)) return BG_ENOMEM;
	CU(cudaMemcpyAsync(c->d_codes.p, Q->codes, ncodes, cudaMemcpyHostToDevice, c->stream));
	CU(cudaMemcpyAsync(c->d_qi.p, c->h_qi.data(), Q->nq * sizeof(QInfo), cudaMemcpyHostToDevice, c->stream));
	if (tasks) CU(cudaMemcpyAsync(c->d_tasks.p, tasks, ntasks * sizeof(bg_task), cudaMemcpyHostToDevice, c->stream));
	c->nq = Q->nq; c->nslots = Q->nslots; c->ntasks = ntasks; c->have_tasks = tasks != nullptr;
	// nominal cell count: 16 * qlen * ClumpLen per task
	uint64_t nominal = 0, fcells = 0;
	if (tasks) {
		for (uint64_t t = 0; t < ntasks; ++t) {
			uint32_t cl = tasks[t].clump - c->first_clump;
			if (cl >= c->num_clumps) continue;
			nominal += 16ull * c->h_qi[tasks[t].query].len * c->clump_len[cl];
			fcells += 16ull * c->h_qi[tasks[t].query].P * c->clump_len[cl];
		}
	} else {
		uint64_t sl = 0, sq = 0, sp = 0;
		for (uint32_t i = 0; i < c->num_clumps; ++i) sl += c->clump_len[i];
		for (uint32_t i = 0; i < Q->nq; ++i) { sq += c->h_qi[i].len; sp += c->h_qi[i].P; }
		nominal = 16ull * sl * sq; fcells = 16ull * sl * sp;
	}
	memset(&c->stats, 0, sizeof(c->stats));
	c->stats.tasks = ntasks; c->stats.nominal_cells = nominal; c->stats.filter_cells = fcells;
	if (!c->surv_cap) c->surv_cap = 1u << 20;
	uint64_t want = std::min<uint64_t>(ntasks * 16, std::max<uint64_t>(c->surv_cap, 4ull * Q->nq + ntasks / 8));
	want = std::max<uint64_t>(want, 1024);
	if (want > c->surv_cap || !c->d_surv.p) c->surv_cap = (uint32_t)std::min<uint64_t>(want, 0xFFFFFFF0ull);
	if (c->d_surv.need(c->surv_cap) || c->d_res.need(c->surv_cap) || c->d_hits.need(c->surv_cap)) return BG_ENOMEM;
	if (!c->d_scratch.p && c->d_scratch.need(1u << 22)) return BG_ENOMEM;
	k_query_prep<<<(Q->nq * 16 + 255) / 256, 256, 0, c->stream>>>(c->d_codes.p, c->d_qi.p, c->d_sterm.p, Q->nq, c->d_peq.p);
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(c->stream));     // the caller's host buffers may be reused after this returns
	c->ran = false; c->by_runs = false;
	return BG_OK;
}

static int run_extend(bg_ctx *c, int mode, const uint16_t *best_in) {
	CU(cudaSetDevice(c->device));
	c->last_mode = mode; c->have_best_in = best_in != nullptr;
	if (best_in) {
		c->last_best_in.assign(best_in, best_in + c->nslots);
		CU(cudaMemcpyAsync(c->d_best16.p, best_in, c->nslots * 2, cudaMemcpyHostToDevice, c->stream));
	}
	k_init_best<<<(c->nslots + 255) / 256, 256, 0, c->stream>>>(c->d_best.p, best_in ? c->d_best16.p : nullptr, c->nslots);
	CU(cudaMemsetAsync(c->d_counters.p, 0, 16, c->stream));
	CU(cudaMemsetAsync(c->d_cells.p, 0, 8, c->stream));
	CU(cudaEventRecord(c->ev[0], c->stream));
	FilterArgs F;
	F.db = c->d_db.p; F.clump_off = c->d_clump_off.p; F.clump_len = c->d_clump_len.p; F.qi = c->d_qi.p;
	F.peq = c->d_peq.p; F.tasks = c->have_tasks ? c->d_tasks.p : nullptr; F.ntasks = c->ntasks; F.nq = c->nq;
	F.first_clump = c->first_clump; F.num_clumps = c->num_clumps; F.surv = c->d_surv.p; F.surv_cap = c->surv_cap;
	F.counters = c->d_counters.p; F.two = 2; F.c16 = 16;
	uint64_t blocks = (c->ntasks + 7) / 8;
	if (blocks > 0x7FFFFFFFull) return fail(BG_EINVAL, "too many tasks for one launch");
	if (blocks) {
		static int variant = getenv("BURST_FILTER_VARIANT") ? atoi(getenv("BURST_FILTER_VARIANT")) : 4;
		if (variant == 1) k_filter<1><<<(unsigned)blocks, 128, 0, c->stream>>>(F);
		else if (variant == 4) k_filter<4><<<(unsigned)blocks, 128, 0, c->stream>>>(F);
		else if (variant == 5) k_filter<5><<<(unsigned)blocks, 128, 0, c->stream>>>(F);
		else if (variant == 3) k_filter<3><<<(unsigned)blocks, 128, 0, c->stream>>>(F);
		else k_filter<2><<<(unsigned)blocks, 128, 0, c->stream>>>(F);
	}
	CU(cudaGetLastError());
	CU(cudaEventRecord(c->ev[1], c->stream));
	ExtendArgs E;
	E.dbw = (const uint32_t *)c->d_db.p; E.clump_off = c->d_clump_off.p; E.clump_len = c->d_clump_len.p;
	E.codes = c->d_codes.p; E.qi = c->d_qi.p; E.tasks = F.tasks; E.nq = c->nq; E.first_clump = c->first_clump;
	E.surv = c->d_surv.p; E.surv_cap = c->surv_cap; E.counters = c->d_counters.p; E.res = c->d_res.p;
	E.best = c->d_best.p; E.Sterm = c->d_sterm.p; E.scratch = c->d_scratch.p; E.scratch_cap = (uint32_t)std::min<size_t>(c->d_scratch.cap, 0xFFFFFFFFu);
	E.band_cells = c->d_cells.p; E.mode = mode;
	const unsigned g = (unsigned)c->sms * 8;
	k_extend<8><<<g, 128, 0, c->stream>>>(E);
	k_extend<16><<<g, 128, 0, c->stream>>>(E);
	k_extend<32><<<g, 128, 0, c->stream>>>(E);
	k_extend<64><<<g, 128, 0, c->stream>>>(E);
	k_extend<0><<<g, 128, 0, c->stream>>>(E);
	CU(cudaGetLastError());
	CU(cudaEventRecord(c->ev[2], c->stream));
	return BG_OK;
}

static int run_select(bg_ctx *c, int mode) {
	CU(cudaSetDevice(c->device));
	k_select<<<(unsigned)c->sms * 4, 256, 0, c->stream>>>(c->d_surv.p, c->d_res.p, c->have_tasks ? c->d_tasks.p : nullptr,
		c->d_qi.p, c->nq, c->d_best.p, c->d_counters.p, c->surv_cap, c->d_hits.p, mode);
	CU(cudaGetLastError());
	CU(cudaEventRecord(c->ev[3], c->stream));
	c->ran = true;
	return BG_OK;
}

extern "C" int bg_batch_run_extend(bg_ctx *c, int mode, const uint16_t *best_in) {
	if (!c || !c->nq) return fail(BG_EINVAL, "bg_batch_run: no batch uploaded");
	return run_extend(c, mode, best_in);
}
extern "C" void *bg_batch_best_device(bg_ctx *c) { return c ? (void *)c->d_best.p : nullptr; }
extern "C" int bg_batch_run_select(bg_ctx *c, int mode) {
	if (!c || !c->nq) return fail(BG_EINVAL, "bg_batch_run: no batch uploaded");
	return run_select(c, mode);
}
extern "C" int bg_batch_run(bg_ctx *c, int mode, const uint16_t *best_in) {
	if (!c || !c->nq) return fail(BG_EINVAL, "bg_batch_run: no batch uploaded");
	int rc = run_extend(c, mode, best_in);
	if (rc) return rc;
	return run_select(c, mode);
}

// Wait for the batch; if the survivor list or the generic-band scratch overflowed, grow and redo.
static int settle(bg_ctx *c) {
	if (!c->ran) return fail(BG_EINVAL, "no batch has been run");
	for (int attempt = 0; attempt < 4; ++attempt) {
		CU(cudaMemcpyAsync(c->h_counters, c->d_counters.p, 16, cudaMemcpyDeviceToHost, c->stream));
		CU(cudaStreamSynchronize(c->stream));
		bool grow_s = c->h_counters[0] > c->surv_cap, grow_g = c->h_counters[1] > c->d_scratch.cap;
		if (!grow_s && !grow_g) return BG_OK;
		if (grow_s) {
			c->surv_cap = c->h_counters[0] + c->h_counters[0] / 4;
			if (c->d_surv.need(c->surv_cap) || c->d_res.need(c->surv_cap) || c->d_hits.need(c->surv_cap)) return BG_ENOMEM;
		}
		if (grow_g && c->d_scratch.need((size_t)c->h_counters[1] + 1024)) return BG_ENOMEM;
		int rc = run_extend(c, c->last_mode, c->have_best_in ? c->last_best_in.data() : nullptr);
		if (rc) return rc;
		rc = run_select(c, c->last_mode);
		if (rc) return rc;
	}
	return fail(BG_EOVERFLOW, "survivor list kept overflowing (%u entries)", c->h_counters[0]);
}

extern "C" int bg_batch_count(bg_ctx *c, uint64_t *nhits) {
	if (!c) return fail(BG_EINVAL, "null ctx");
	CU(cudaSetDevice(c->device));
	int rc = settle(c); if (rc) return rc;
	if (nhits) *nhits = c->h_counters[2];
	return BG_OK;
}

extern "C" int bg_batch_download(bg_ctx *c, bg_hit *hits, uint64_t cap, uint16_t *best_out) {
	if (!c) return fail(BG_EINVAL, "null ctx");
	CU(cudaSetDevice(c->device));
	int rc = settle(c); if (rc) return rc;
	uint64_t n = c->h_counters[2];
	if (hits) {
		if (cap < n) return fail(BG_EINVAL, "bg_batch_download: %llu hits, room for %llu", (unsigned long long)n, (unsigned long long)cap);
		CU(cudaMemcpyAsync(hits, c->d_hits.p, n * sizeof(bg_hit), cudaMemcpyDeviceToHost, c->stream));
	}
	std::vector<uint32_t> b32;
	if (best_out) {
		b32.resize(c->nslots);
		CU(cudaMemcpyAsync(b32.data(), c->d_best.p, c->nslots * 4, cudaMemcpyDeviceToHost, c->stream));
	}
	CU(cudaStreamSynchronize(c->stream));
	if (best_out) for (uint32_t i = 0; i < c->nslots; ++i) best_out[i] = (uint16_t)std::min<uint32_t>(b32[i], 0xFFFF);
	if (hits) std::sort(hits, hits + n, [](const bg_hit &a, const bg_hit &b) { return a.task != b.task ? a.task < b.task : a.lane < b.lane; });
	if (hits && c->by_runs) for (uint64_t i = 0; i < n; ++i) hits[i].task = c->run_key[hits[i].task];
	return BG_OK;
}

extern "C" int bg_batch_stats(bg_ctx *c, bg_stats *out) {
	if (!c || !out) return fail(BG_EINVAL, "null argument");
	CU(cudaSetDevice(c->device));
	int rc = settle(c); if (rc) return rc;
	unsigned long long cells = 0;
	CU(cudaMemcpy(&cells, c->d_cells.p, 8, cudaMemcpyDeviceToHost));
	c->stats.survivors = c->h_counters[0]; c->stats.hits = c->h_counters[2]; c->stats.band_cells = cells;
	cudaEventElapsedTime(&c->stats.ms_filter, c->ev[0], c->ev[1]);
	cudaEventElapsedTime(&c->stats.ms_extend, c->ev[1], c->ev[2]);
	cudaEventElapsedTime(&c->stats.ms_select, c->ev[2], c->ev[3]);
	*out = c->stats;
	return BG_OK;
}

extern "C" int bg_align_batch(bg_ctx *c, const bg_queries *Q, const bg_task *tasks, uint64_t ntasks, int mode,
		uint16_t *best_inout, bg_hit **hits, uint64_t *nhits) {
	if (!hits || !nhits) return fail(BG_EINVAL, "bg_align_batch: null output");
	int rc = bg_batch_upload(c, Q, tasks, ntasks); if (rc) return rc;
	rc = bg_batch_run(c, mode, best_inout); if (rc) return rc;
	uint64_t n = 0;
	rc = bg_batch_count(c, &n); if (rc) return rc;
	bg_hit *h = (bg_hit *)malloc((n ? n : 1) * sizeof(bg_hit));
	if (!h) return fail(BG_ENOMEM, "malloc hits");
	rc = bg_batch_download(c, h, n, best_inout);
	if (rc) { free(h); return rc; }
	*hits = h; *nhits = n;
	return BG_OK;
}

// Run lists (bg_run): expanded to tasks here; hits come back keyed run * BG_RUN_MAX + i.
static int expand_runs(const bg_queries *Q, const bg_run *runs, uint64_t nruns, std::vector<bg_task> &tasks, std::vector<uint32_t> &key) {
	if (!Q || !runs) return fail(BG_EINVAL, "bg_batch_upload_runs: null argument");
	if (nruns >= (1ull << 28)) return fail(BG_EINVAL, "bg_batch_upload_runs: %llu runs in one batch (limit 2^28-1)", (unsigned long long)nruns);
	tasks.clear(); key.clear();
	for (uint64_t r = 0; r < nruns; ++r) {
		if (!runs[r].nq || runs[r].nq > BG_RUN_MAX || (uint64_t)runs[r].query0 + runs[r].nq > Q->nq)
			return fail(BG_EINVAL, "bg_batch_upload_runs: run %llu (query0 %u, nq %u) is malformed", (unsigned long long)r, runs[r].query0, runs[r].nq);
		for (uint32_t i = 0; i < runs[r].nq; ++i) { tasks.push_back(bg_task{runs[r].query0 + i, runs[r].clump}); key.push_back((uint32_t)(r * BG_RUN_MAX + i)); }
	}
	return BG_OK;
}
extern "C" int bg_batch_upload_runs(bg_ctx *c, const bg_queries *Q, const bg_run *runs, uint64_t nruns) {
	if (!c) return fail(BG_EINVAL, "null ctx");
	std::vector<bg_task> tasks;
	int rc = expand_runs(Q, runs, nruns, tasks, c->run_key); if (rc) return rc;
	rc = bg_batch_upload(c, Q, tasks.data(), tasks.size());
	c->by_runs = rc == BG_OK;
	return rc;
}
extern "C" int bg_align_runs(bg_ctx *c, const bg_queries *Q, const bg_run *runs, uint64_t nruns, int mode,
		uint16_t *best_inout, bg_hit **hits, uint64_t *nhits) {
	if (!hits || !nhits) return fail(BG_EINVAL, "bg_align_runs: null output");
	int rc = bg_batch_upload_runs(c, Q, runs, nruns); if (rc) return rc;
	rc = bg_batch_run(c, mode, best_inout); if (rc) return rc;
	uint64_t n = 0;
	rc = bg_batch_count(c, &n); if (rc) return rc;
	bg_hit *h = (bg_hit *)malloc((n ? n : 1) * sizeof(bg_hit));
	if (!h) return fail(BG_ENOMEM, "malloc hits");
	rc = bg_batch_download(c, h, n, best_inout);
	if (rc) { free(h); return rc; }
	*hits = h; *nhits = n;
	return BG_OK;
}
extern "C" void bg_free_hits(bg_hit *h) { free(h); }
