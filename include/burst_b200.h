/* burst_b200.h -- C ABI of the B200 alignment engine (libburst_b200.so).
 *
 * This is the drop-in boundary for BURST's alignment hot path.  The reference has no
 * plugin/FFI layer: the path is three static inline C functions called from one place each
 * (SURVEY.md 8b).  Their per-call, latency-shaped signatures
 *
 *   uint32_t aded_mat16L(DualCoil *ref, char *query, uint32_t rwidth, uint32_t qlen, uint32_t width,
 *                        uint32_t minlen, DualCoil *Matrix, DualCoil *profile, uint32_t maxED,
 *                        uint32_t startQ, uint32_t *LoBound, uint32_t *HiBound, DualCoil *MinA)
 *                                                                     burst.c:1106-1107
 *   uint32_t aded_mat16 / aded_xalpha (same minus minlen)             burst.c:1097-1101
 *   void reScoreM_mat16 / reScoreM_xalpha(DualCoil *ref, char *query, uint32_t rwidth, uint32_t qlen,
 *                        uint32_t width, DualCoil *Matrix, DualCoil *Shifts, DualCoil *ShiftR,
 *                        uint32_t maxED, DualCoil *profile, MetaPack *M16)   burst.c:890-896
 *
 * are replaced by ONE batched call over a list of (query, clump) tasks -- the pairs the
 * reference's drivers enumerate at burst.c:4137/4157 (accelerated) and 4344/4365 (all-vs-all)
 * -- returning, per task and lane, what the reference would have pushed as a ResultPod
 * (burst.c:3999-4004, 4228-4238): the lanes whose edit distance equals the per-query minimum
 * (BEST / ALLPATHS / CAPITALIST) or all lanes within budget (FORAGE), each with the pass-2
 * integers of MetaPack (burst.c:222-226).  The float identity is derived on the host from
 * these integers exactly as burst.c:844-860 does.
 *
 * Conventions: plain pointers and sizes; the caller owns every buffer it passes; every function
 * returns 0 on success and a non-zero BG_E* code on failure, with a message available from
 * bg_last_error() (the reference prints and exit()s; the CLI layer maps codes to its exit
 * statuses).  A context is bound to one CUDA device and one host thread at
 * a time.  There is no CPU fallback behind this ABI: without a CUDA device bg_init fails.
 */
#ifndef BURST_B200_H
#define BURST_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct bg_ctx bg_ctx;

enum { BG_OK = 0, BG_EINVAL = 1, BG_ECUDA = 3, BG_EOVERFLOW = 4, BG_ENOMEM = 5 };

/* Which lanes are reported (burst.c:4219-4224): */
enum { BG_MODE_MIN = 0,   /* lanes at the per-slot minimum: BEST, ALLPATHS, CAPITALIST */
       BG_MODE_ALL = 1 }; /* every lane within budget: FORAGE */

/* One DP task = one call pair aded_*() + reScoreM_*() of the reference: query j vs clump ri. */
typedef struct { uint32_t query, clump; } bg_task;

/* One clump visit of the reference's bunch loop (burst.c:4137-4157: unpack clump ri once, then
 * "for each query j in the bunch"): the nq <= BG_RUN_MAX consecutive queries query0 .. query0+nq-1
 * of the batch against one clump.  A run is the unit the GPU schedules (one warp scans the clump
 * once for all its queries); a bunch x candidate-list driver emits runs directly, 1/16th the bytes
 * of the equivalent bg_task list.  Hits of a run-list batch carry task = run * BG_RUN_MAX + (query - query0). */
#define BG_RUN_MAX 16
typedef struct { uint32_t clump, query0, nq; } bg_run;

/* One reported lane = one ResultPod (burst.c:3999-4004) minus the host-only fields.
 * refIx = tasks[task].clump * 16 + lane (burst.c:4234); ed = mismatches; gap_q = numGapQ
 * (shift), gap_r = numGapR (shiftR), final_pos = 1-based end column in the clump. */
typedef struct { uint32_t task; uint8_t lane, ed, gap_q, gap_r; uint32_t final_pos; } bg_hit;

/* A batch of queries in BURST's translated form (burst.c:3005-3011): code bytes 0..15.
 * budget[i] = Emac the reference would start query i with (ShrBin.ed, burst.c:3069-3081),
 * slot[i]   = index of the running-minimum cell the query shares with its reverse complement
 *             (Ub->six, burst.c:4158, 4218); slot[i] < nslots. */
enum { BG_Q_PACKED4 = 1 };  /* bg_queries.flags: `codes` holds two bases per byte (even base in the low nibble, the .edx
                             * clump convention, burst.c:2810-2824) and offset[] counts bases, so a query may start on
                             * either nibble of a byte.  Halves the bytes that cross PCIe; results are identical. */
typedef struct {
	const uint8_t  *codes;
	const uint64_t *offset;     /* nq + 1 entries into codes (in bases) */
	const uint16_t *budget;     /* nq, each <= 254 */
	const uint32_t *slot;       /* nq */
	uint32_t nq, nslots;
	uint32_t flags;             /* 0, or BG_Q_PACKED4 */
} bg_queries;

typedef struct {
	uint64_t tasks;             /* (query, clump) pairs evaluated */
	uint64_t nominal_cells;     /* sum over tasks of 16 * qlen * ClumpLen */
	uint64_t filter_cells;      /* DP cells covered by the Myers prefix filter (k_filter) */
	uint64_t seed_steps;        /* (lane, column) positions streamed by the pigeonhole seed filter (k_seed), once per run */
	uint64_t survivors;         /* (task, lane) pairs handed to the banded pass */
	uint64_t band_cells;        /* DP cells updated by the banded pass (x3 values each) */
	uint64_t hits;              /* lanes reported */
	uint32_t seed_queries;      /* queries of the batch taken by k_seed (the rest go through k_filter) */
	uint32_t seed_stride, seed_window;       /* window layout chosen for the batch: probe every `stride` columns, `window` bases */
	uint32_t seed_words;                     /* words of the per-warp window filter */
	float ms_filter, ms_extend, ms_select;   /* device time of the last bg_batch_run */
} bg_stats;

/* ---- lifetime ---- */
int  bg_init(int device, bg_ctx **ctx);
void bg_free(bg_ctx *ctx);
const char *bg_last_error(void);
/* Run every kernel and copy of this context on an existing CUDA stream (cudaStream_t). */
int  bg_set_stream(bg_ctx *ctx, void *cuda_stream);

/* Tuning knobs (results never depend on them). */
enum { BG_PARAM_SEED_FILTER = 1,     /* 1 (default): pigeonhole seed filter where the batch allows; 0: Myers prefix filter only */
       BG_PARAM_SEED_CHUNK  = 2,     /* consecutive runs handled by one warp of the seed filter (default 8) */
       BG_PARAM_SEED_WORDS  = 3,     /* 32-bit words of the per-warp window filter, power of two 128..8192 (0 = sized from the batch) */
       BG_PARAM_SEED_STAGE  = 4,     /* 1: clumps reach the seed filter through bulk copies (TMA, cp.async.bulk + mbarrier) into shared memory,
                                        one run ahead; 0 (default): direct 128-bit loads, which leave room for more resident blocks */
       BG_PARAM_PIPE_SLICES = 5 };   /* slices the one-call run-list path cuts a large batch into so that host->device copies overlap the kernels
                                        (default 4; 0 or 1 = one upload, then run) */
enum { BG_PARAM_SEED_GROUPS = 8 };   /* 16-thread groups (runs per round) per block of the seed filter: 2, 4 or 8; 0 (default) = 4 when the tables allow */
enum { BG_PARAM_PIPE_MIN_RUNS = 6,   /* fewest runs worth a slice (default 4096): lists shorter than two slices take the single-batch path */
       BG_PARAM_PIPE_RATIO = 7 };    /* size of each slice in percent of the one before; 0 (default) = 100 for byte codes (copy-bound), 140 for
                                        BG_Q_PACKED4 (kernel-bound: a short first copy, later copies hide behind the kernels) */
enum { BG_PARAM_SEED_IMPL = 9,       /* 1 (default): warp-per-bunch seed filter (private window table, no block barriers); 0: the block form */
       BG_PARAM_SEED_NCH = 10,       /* 32-column chunks per staged item of the warp form: 4..8; 0 (default) = the size that wastes the fewest probes on the loaded database */
       BG_PARAM_SEED_LBITS = 11,     /* log2 of the bits in a warp's window filter, 10..20 (0 = sized from the batch) */
       BG_PARAM_SEED_FB = 12,        /* bits set per window in that filter: 2 (default) or 1 */
       BG_PARAM_SEED_HSLOTS = 13,    /* buckets of a warp's chained window table, a power of two 64..4096 (0 = sized from the batch) */
       BG_PARAM_SEED_VMODE = 14 };   /* how the warp form verifies flagged words: 0 = queue + helper lanes, 1 = one pass per flagged lane with warp reductions */
int  bg_set_param(bg_ctx *ctx, int what, int value);

/* Page-locked host memory for the arrays handed to bg_align_runs_into() (queries, runs, hit buffer): makes the library's
 * host<->device copies asynchronous, so that they overlap its kernels.  bg_host_free(NULL) is a no-op. */
void *bg_host_alloc(uint64_t bytes);
void  bg_host_free(void *p);

/* ---- scoring: the 16x16 table the reference builds in setScore() (burst.c:1309-1328),
 * S[q*16+r] in {0,1,255}; call before aligning (default: Z=1 table). */
int  bg_set_scoring(bg_ctx *ctx, const uint8_t S[256]);
/* Fill S with the reference's table for Z (1 = penalise N, the default; 0 = -y). */
void bg_default_scoring(int z, uint8_t S[256]);

/* ---- database: clumps in the .edx on-disk clump layout (burst.c:2810-2824): clump i is
 * ceil(clump_len[i]/2) vectors of 16 bytes, byte k of a vector = lane k, low nibble = position
 * 2v, high nibble = position 2v+1; clumps are contiguous in `packed`.  first_clump is the
 * global id of packed clump 0 (non-zero when this device holds one shard of a larger DB;
 * tasks naming clumps outside [first_clump, first_clump+num_clumps) are skipped). */
int  bg_load_db(bg_ctx *ctx, const uint8_t *packed, const uint32_t *clump_len,
                uint32_t num_clumps, uint32_t first_clump);

/* ---- one batch, three steps (resident form used by bench.py's kernel-only timing) ---- */
/* Host -> device copy of queries and tasks.  tasks == NULL means all-vs-all in the
 * reference's fallback order (burst.c:4344, 4365): task t = clump (t / nq), query (t % nq). */
int  bg_batch_upload(bg_ctx *ctx, const bg_queries *q, const bg_task *tasks, uint64_t ntasks);
/* The same with the task list given as runs (see bg_run). */
int  bg_batch_upload_runs(bg_ctx *ctx, const bg_queries *q, const bg_run *runs, uint64_t nruns);
/* best_in: nslots running minima carried in from earlier batches (NULL = none). */
int  bg_batch_run(bg_ctx *ctx, int mode, const uint16_t *best_in);
/* Split form of bg_batch_run for a reference-sharded DB: filter+extend, then an external
 * all-reduce(MIN) over the device array bg_batch_best_device() (nslots x uint32), then select. */
int  bg_batch_run_extend(bg_ctx *ctx, int mode, const uint16_t *best_in);
void *bg_batch_best_device(bg_ctx *ctx);
/* The CUDA stream (cudaStream_t) the context's kernels run on: enqueue the all-reduce there and no host synchronisation is needed
 * between bg_batch_run_extend, the all-reduce and bg_batch_run_select. */
void *bg_stream(bg_ctx *ctx);
/* Test hook: the next batch starts with a survivor list of exactly `cap` entries (the engine grows it and redoes the work on overflow). */
int  bg_set_surv_cap(bg_ctx *ctx, uint32_t cap);
int  bg_batch_run_select(bg_ctx *ctx, int mode);
/* Device -> host: number of hits, then the hits sorted by (task, lane), and the per-slot
 * minima (0xFFFF = no lane within budget).  Either output pointer may be NULL. */
int  bg_batch_count(bg_ctx *ctx, uint64_t *nhits);
int  bg_batch_download(bg_ctx *ctx, bg_hit *hits, uint64_t cap, uint16_t *best_out);
int  bg_batch_stats(bg_ctx *ctx, bg_stats *out);

/* ---- compact strand batches: the form of the accelerated path with -fr (burst.c:3087-3109, 4077-4157) that moves the fewest bytes.
 * Every READ crosses the bus once, 4 or 2 bits per base; the device derives both strands.  The queries of the batch are the
 * STRANDS in the order the host sorted them (UniBins order): strand[q] = read index | (1 << 31 if reverse complement); the slot of
 * a strand (its running minimum, shared by both strands, burst.c:4218) is its read index, so best_inout has nreads entries.
 * Work comes as the reference's bunch -> candidate lists: bunch b = strands b*qbunch .. (+qbunch), its candidates
 * cand[cand_off[b] .. cand_off[b+1]) in visiting order.  Run r of the batch = candidate r; hits carry task = r * BG_RUN_MAX + (strand - b*qbunch). */
enum { BG_R_PACKED4 = 1,    /* reads: concatenated code nibbles, base i of the stream in nibble i & 1 of byte i >> 1 (low first) */
       BG_R_PACKED2 = 2 };  /* reads: A C G T only, code - 1 in bits 2(i & 3) .. of byte i >> 2 */
typedef struct {
	const uint8_t  *reads;      /* packed stream of all reads, back to back (no per-read alignment) */
	const uint16_t *len;        /* nreads lengths in bases (>= 1) */
	const uint16_t *budget;     /* nreads budgets (ShrBins[].ed, <= 254) */
	const uint32_t *strand;     /* nq strands in sorted order */
	uint32_t nreads, nq, flags; /* flags: BG_R_PACKED4 or BG_R_PACKED2 */
} bg_reads;
int  bg_align_bunches_into(bg_ctx *ctx, const bg_reads *reads, uint32_t qbunch, const uint32_t *cand_off, const uint32_t *cand, uint32_t nbunch,
                           int mode, uint16_t *best_inout, bg_hit *hits, uint64_t cap, uint64_t *nhits);

/* ---- candidate generation on the device (the .acx lookup of burst.c:4085-4133): load the accelerator once, then hand over strand
 * batches WITHOUT candidate lists.  lens: the 4^word_len posting-list lengths as stored in the file (burst.c:3504-3506); postings: the
 * packed lists that follow them (small format: two 20-bit clump ids in 5 bytes, an odd last one in 3; big: 3 bytes each, burst.c:3512-3527);
 * bad: the always-visited clumps (burst.c:3530). */
int  bg_load_acx(bg_ctx *ctx, const uint32_t *lens, const uint8_t *postings, uint64_t post_bytes, int word_len, int big,
                 const uint32_t *bad, uint32_t nbad);
/* A hit of a device-generated run list: the strand (index into reads->strand) and clump instead of a task number. */
typedef struct { uint32_t query, clump; uint8_t lane, ed, gap_q, gap_r; uint32_t final_pos; } bg_xhit;
/* Strand batch in, hits out: candidates per bunch of `qbunch` strands exactly as the reference picks and orders them (count above the
 * bunch threshold, descending count, ties in first-touch order; per-query skip; BadList unless skip_bad), then the alignment.  Reads must
 * be BG_R_PACKED2 (plain bases: the reference expands ambiguous query bases into all variants, which stays a host job).  Hits arrive
 * grouped by bunch, inside a bunch in the reference's visiting order (candidate, query, lane). */
int  bg_search_bunches_into(bg_ctx *ctx, const bg_reads *reads, uint32_t qbunch, int heuristic, int skip_bad, int mode,
                            uint16_t *best_inout, bg_xhit *hits, uint64_t cap, uint64_t *nhits);

/* ---- the one-call form the host driver uses: upload + run + download.  *hits is
 * malloc()ed by the library (free with bg_free_hits). */
int  bg_align_batch(bg_ctx *ctx, const bg_queries *q, const bg_task *tasks, uint64_t ntasks,
                    int mode, uint16_t *best_inout, bg_hit **hits, uint64_t *nhits);
int  bg_align_runs(bg_ctx *ctx, const bg_queries *q, const bg_run *runs, uint64_t nruns,
                   int mode, uint16_t *best_inout, bg_hit **hits, uint64_t *nhits);
void bg_free_hits(bg_hit *hits);
/* The same with the hits written to a caller-owned buffer of `cap` entries (pinned host memory makes the copy
 * asynchronous): no allocation per call.  BG_EOVERFLOW with *nhits = the number needed when cap is too small. */
int  bg_align_runs_into(bg_ctx *ctx, const bg_queries *q, const bg_run *runs, uint64_t nruns,
                        int mode, uint16_t *best_inout, bg_hit *hits, uint64_t cap, uint64_t *nhits);

/* Two batches in flight on one GPU.  `ctx` (same device as `src`) uses the database -- and the accelerator, when one is loaded -- that
 * `src` already holds in HBM instead of a copy of its own; everything else (streams, batch buffers, survivor lists) stays private, so one
 * host thread per context can run bg_align_*_into() / bg_search_bunches_into() concurrently: the host<->device copies of one batch
 * travel behind the kernels of the other.  This is what the reference's thread team does with its per-thread scratch over one shared
 * database (burst.c:4050-4077).  In bg_align_bunches_into() the filter + sweep kernels of the sharing contexts are queued one batch behind
 * the other (a shared event chain), so that two threads that start together do not fall into lockstep; the other entry points run their
 * kernels concurrently.  `src` must not be freed, nor load another database, before `ctx` is freed. */
int  bg_share_db(bg_ctx *ctx, bg_ctx *src);

#ifdef __cplusplus
}
#endif
#endif
