/*
 * nwb.h -- C ABI of the B200-native Needleman-Wunsch score-table fill.
 *
 * This is the drop-in boundary for ONE path of skotchandsoda/needleman-wunsch:
 * the score-table fill.  Each entry point cites the reference interface it
 * replaces (paths are into the reference repository).  C99-includable,
 * extern "C", plain pointers and sizes only.
 *
 * Orientation (reference naming, computation.c:158-160): the TOP string s1
 * (length A) indexes columns i in [0,A]; the SIDE string s2 (length B) indexes
 * rows j in [0,B]; the table is M x N = (A+1) x (B+1); the reference stores
 * cells[i][j].
 *
 * There is no CPU fallback: every fill runs hand-written sm_100a kernels and
 * fails with NWB_ERR_NO_DEVICE / NWB_ERR_CUDA when no usable GPU is present.
 *
 * ARROW TABLE LAYOUT (replaces walk_table_cell_t{diag,left,up}, walk-table.h:48-57,
 * 32 B/cell, by 4 bits/cell):
 *   - interior cells only (i in [1,A], j in [1,B]); the borders are implied
 *     (row 0: LEFT, column 0: UP, (0,0): none -- computation.c:97-124);
 *   - row-major: row j starts at byte (j-1)*pitch, pitch = nwb_arrow_pitch()
 *     (a multiple of 16 bytes, >= ceil(A/2));
 *   - cell (i,j) is the 4-bit code in byte (j-1)*pitch + (i-1)/2, low nibble
 *     when (i-1) is even: bit0 = DIAG, bit1 = LEFT, bit2 = UP, bit3 = 0.
 *     Several bits are set when candidates tie (needleman-wunsch.c:485-503
 *     records EVERY candidate equal to the maximum).
 *   - bytes/nibbles beyond column A inside the pitch are unspecified.
 *   The reference's `match` flag (score-table.h:60, written at
 *   needleman-wunsch.c:432-438, never read back) is derived from the strings by
 *   nwb_arrows() and returned as NWB_MATCH.
 */
#ifndef NWB_H
#define NWB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- arrow code bits returned by nwb_arrows() ---------------------------- */
#define NWB_DIAG 1u  /* walk_table_cell_t.diag  */
#define NWB_LEFT 2u  /* walk_table_cell_t.left  */
#define NWB_UP 4u    /* walk_table_cell_t.up    */
#define NWB_MATCH 8u /* score_table_cell_t.match */

/* ---- flags --------------------------------------------------------------- */
/* Materialise the int32 score matrix (needed by the -t printer,
 * print-table.c:120-157, and by score-level verification).  O(A*B*4) bytes. */
#define NWB_WANT_SCORES 0x01u
/* Fuse the 64-bit optimal-alignment count into the fill (-s; replaces the
 * exponential enumeration behind get_solution_count(), computation.c:249). */
#define NWB_WANT_COUNT 0x02u
/* Copy the arrow table to host memory so nwb_arrows()/nwb_arrow_rows() work
 * (the walk needs it: needleman-wunsch.c:667 `!q || l || s || t`). */
#define NWB_WANT_ARROWS_HOST 0x04u
/* Track max |score| over interior cells (the tflag rule,
 * needleman-wunsch.c:538-541; score_table_t.greatest_abs_val). */
#define NWB_TRACK_ABS 0x08u
/* Always use the general int32 kernel (any m,k,d) instead of the packed
 * 16x2 difference kernel that is selected automatically for small scores. */
#define NWB_FORCE_GENERAL 0x10u
/* Verification aid: materialise the full uint64 count matrix (small inputs). */
#define NWB_WANT_COUNT_MATRIX 0x20u
/* Do not produce the branch counter (walk_table_t.branch_count, which the
 * reference only prints from debug builds, needleman-wunsch.c:624).  The packed
 * kernel obtains it with a second pass over the arrow table; this skips it. */
#define NWB_NO_BRANCH_COUNT 0x40u
/* Verification aid: obtain the count with the dense forward sweep and also return the digests of the count
 * matrix's bottom row and last column (nwb_summary.lastrow_count_digest / lastcol_count_digest).  Implies
 * NWB_WANT_COUNT. */
#define NWB_WANT_COUNT_DIGEST 0x80u
/* nwb_fill()/nwb_fill_on(): compute the arrow-table digest on the device before returning (nwb_arrow_digest()). */
#define NWB_WANT_DIGEST 0x100u
/* nwb_plan_create(): this plan is one of several on the same GPU that work through a QUEUE of fills (fill e on plan
 * e mod n, each plan on its own stream).  The single-pair kernel is then launched so that consecutive fills overlap:
 * its blocks draw tickets and sweep adjacent strips, need no co-residency, and the next fill's blocks move onto SMs as
 * this fill's leave them.  One fill alone is ~5 % slower this way; a queue runs at the rate at which the SMs sweep
 * strips instead of the rate of one strip-to-strip wavefront. */
#define NWB_QUEUE 0x200u

/* ---- error codes ---------------------------------------------------------- */
#define NWB_OK 0
#define NWB_ERR_INVALID (-1)     /* bad argument                               */
#define NWB_ERR_NOMEM (-2)       /* host or device allocation failed           */
#define NWB_ERR_CUDA (-3)        /* a CUDA call or kernel failed               */
#define NWB_ERR_NO_DEVICE (-4)   /* no CUDA device: there is no CPU fallback   */
#define NWB_ERR_UNSUPPORTED (-5) /* flag combination not available             */

const char *nwb_strerror(int err);
/* Explicit overrides for tests and measurements; the library reads NO environment variables.  Every key
 * selects between kernels that produce identical results ("pk_k", "pk_r", "pk_warps", "pk_hx", "count_mode"
 * 0 auto / 1 fused into the fill / 2 dense sweep after the fill / 3 dense sweep trailing the fill on a second
 * stream, "cnt_cpl", "batch_bx", "batch_cx", "batch_bp", "bp_warps", "bp_aligned", "batch_lcount", "lc_warps", "bcnt_chain", "cx_warps",
 * "hx_spb" 1..3 = the single-pair kernel in queue mode -- see NWB_QUEUE -- with that many adjacent strips per block; 0 or -1 = automatic,
 * see nwb_api.cu) except "watchdog_ms" (how long a device-side wait may see no progress before the fill
 * fails with NWB_ERR_CUDA; default 4000) and "inject_fault" (test only: 1 makes the next fills lose a
 * strip's boundary stream so that the watchdog path can be exercised -- the fill FAILS, it never returns a
 * wrong result).  Process-wide; set before creating the plan/batch that should see it.  Unknown key:
 * NWB_ERR_INVALID.  nwb_tune_reset() restores the defaults. */
int nwb_tune(const char *key, int value);
void nwb_tune_reset(void);
/* Text of the last CUDA error seen by this thread ("" if none). */
const char *nwb_last_cuda_error(void);
/* Page-locked ("pinned") host memory for input strings: host-to-device copies from it run at full link speed and
 * overlap with the kernels (nwb_batch_refill); any other host memory works too, only slower.  NULL on failure. */
void *nwb_host_alloc(size_t bytes);
void nwb_host_free(void *p);
/* Number of usable CUDA devices (0 if none), ABI version. */
int nwb_device_count(void);
int nwb_abi_version(void);

/* ========================================================================== *
 * 1. One pair, host buffers -- replaces the call
 *        compute_table_scores(C)                 needleman-wunsch.c:583-626
 *    together with the table set-up it depends on
 *        init_computation_tables()               computation.c:75-125
 *        alloc_score_table()/alloc_walk_table()  score-table.c:60, walk-table.c:58
 *    at the call site needleman_wunsch(), needleman-wunsch.c:654-689.
 * ========================================================================== */
typedef struct nwb_table nwb_table; /* opaque; owns device + host buffers */

/* Blocking.  top/side need not be NUL-terminated; arbitrary bytes, compared
 * for equality (needleman-wunsch.c:432).  m = match bonus, k = mismatch
 * penalty, d = indel penalty, exactly the reference's operands (may be 0 or
 * negative).  On success *out owns the results until nwb_free(). */
int nwb_fill(const char *top, int top_len, const char *side, int side_len,
             int m, int k, int d, unsigned flags, nwb_table **out);
/* As nwb_fill, on CUDA device `device`, splitting the table into column
 * strips over `num_gpus` consecutive devices starting at `device` (boundary
 * columns are pipelined device-to-device over NVLink peer memory). */
int nwb_fill_on(const char *top, int top_len, const char *side, int side_len,
                int m, int k, int d, unsigned flags, int device, int num_gpus,
                nwb_table **out);
/* replaces free_computation(), computation.c:200-214 */
void nwb_free(nwb_table *t);
/* nwb_fill()/nwb_fill_on() keep the device workspace (streams, buffers) of the last fills per (device, number of
 * GPUs) and reuse it when it is large enough, instead of allocating and freeing it on every call.  This releases
 * whatever is cached (e.g. the 5 GB arrow table of a 100k x 100k fill).  nwb_tune("plan_cache", 0) disables it. */
void nwb_cache_clear(void);

/* score_table_t.M-1 / N-1 (score-table.h:70-71) */
int nwb_top_len(const nwb_table *t);
int nwb_side_len(const nwb_table *t);
/* cells[M-1][N-1].score, as printed by print_summary(), computation.c:279-280 */
int32_t nwb_opt_score(const nwb_table *t);
/* Number of optimal alignments mod 2^64 (NWB_WANT_COUNT).  Its low 32 bits are
 * what the reference's `unsigned int solution_count` holds (computation.h:65)
 * and prints with %d (computation.c:277). */
uint64_t nwb_count_u64(const nwb_table *t);
/* get_branch_count(), walk-table.c:133-147 (mod 2^32) */
uint32_t nwb_branch_count(const nwb_table *t);
/* score_table_t.greatest_abs_val under tflag (NWB_TRACK_ABS), else 0 */
int32_t nwb_greatest_abs_interior(const nwb_table *t);
/* cells[i][j].score for any 0<=i<=A, 0<=j<=B (NWB_WANT_SCORES; borders by the
 * init formula).  Returns INT32_MIN when scores were not requested. */
int32_t nwb_score(const nwb_table *t, int i, int j);
/* NWB_DIAG|NWB_LEFT|NWB_UP|NWB_MATCH of cell (i,j), borders included
 * (NWB_WANT_ARROWS_HOST).  Returns 0xFFFFFFFF when arrows are not on host. */
unsigned nwb_arrows(const nwb_table *t, int i, int j);
/* Raw host view of the arrow table (layout above); NULL if not on host. */
const uint8_t *nwb_arrow_rows(const nwb_table *t);
size_t nwb_arrow_pitch(const nwb_table *t);
/* Verification aid (NWB_WANT_COUNT_MATRIX): count of cell (i,j), borders = 1. */
uint64_t nwb_count_at(const nwb_table *t, int i, int j);
/* Raw host views of the INTERIOR score / count matrices: element (i,j), i,j>=1,
 * is rows[(j-1)*(*pitch_elems) + (i-1)].  NULL when not requested. */
const int32_t *nwb_score_rows(const nwb_table *t, size_t *pitch_elems);
const uint64_t *nwb_count_rows(const nwb_table *t, size_t *pitch_elems);
/* Device time of the fill kernel(s) of this table in ms (CUDA events on the
 * launching stream), and which kernel ran: 0 = general int32, 1 = packed 16x2. */
float nwb_kernel_ms(const nwb_table *t);
int nwb_kernel_kind(const nwb_table *t);
/* Parity aid (NWB_WANT_DIGEST): an order-independent 64-bit digest of the whole arrow table, computed on the
 * device: the sum mod 2^64 over rows j = 1..B and 32-bit words w of the table of
 * mix64((j << 32) | w, word(j,w) & 0x77777777 with the cells beyond column A zeroed), mix64(p, x) = the
 * SplitMix64 finaliser of (p + 1) * 0x9E3779B97F4A7C15 + x.  The CPU oracle computes the same sum from the
 * reference's definition of every cell (needleman-wunsch.c:485-503). */
int nwb_arrow_digest(const nwb_table *t, uint64_t *digest);

/* ========================================================================== *
 * 2. Device-resident plan -- the same fill with the strings already in HBM and
 *    the results left there; used for repeated fills and for measurement.
 *    (No reference counterpart: the reference has no device boundary.)
 * ========================================================================== */
typedef struct nwb_plan nwb_plan;

typedef struct nwb_summary {
    int32_t opt_score;
    uint32_t branch_count;
    int32_t greatest_abs;
    int32_t kernel_kind;
    uint64_t count;
    /* packed kernel only: this rank's share of sum_i u(i,B).  For a single-rank
     * plan opt_score is already final; in a strip group the true score is
     * sum over ranks of partial_r, minus d*(A+B). */
    int64_t partial_r;
    /* How the count was obtained: 0 none, 1 fused into the fill, 2 dense forward sweep over the arrow codes,
     * 3 sparse backward sweep from (A,B) over the cells on optimal paths, 4 = 3 gave up (live band wider than
     * its window) and the dense sweep produced the count.  count_rows: rows the sparse sweep visited. */
    int32_t count_path;
    uint32_t count_rows;
    /* NWB_WANT_COUNT_DIGEST: sum_i mix64(i, cnt(i,B)) over this rank's columns i, and sum_j mix64(j, cnt(A,j))
     * (non-zero only on the rank that owns column A); in a strip group the ranks' values add up. */
    uint64_t lastrow_count_digest;
    uint64_t lastcol_count_digest;
} nwb_summary;
/* The summary of a finished nwb_fill()/nwb_fill_on() table (strip groups already combined). */
int nwb_table_summary(const nwb_table *t, nwb_summary *out);

/* Allocate device workspace for fills up to max_top x max_side on `device`.
 * strip_rank/strip_world != (0,1) makes this plan one rank of a column-strip
 * group: the table's 256-column strips are dealt out in contiguous runs of ceil(n_strips / world) strips, rank r
 * taking run r (nwb_strip_partition() below is the same rule; trailing ranks can be empty). */
int nwb_plan_create(int max_top, int max_side, unsigned flags, int device,
                    int strip_rank, int strip_world, nwb_plan **out);
void nwb_plan_destroy(nwb_plan *p);
/* Host -> device copy of the two strings (synchronous on the plan's stream). */
int nwb_plan_upload(nwb_plan *p, const char *top, int top_len, const char *side, int side_len);
/* Launch the fill on `stream` (a cudaStream_t; NULL = the plan's own stream).
 * Asynchronous; inputs and outputs stay on the device. */
int nwb_plan_run(nwb_plan *p, int m, int k, int d, void *stream);
/* Wait for the plan's work and fetch the summary (device -> host).  Returns NWB_ERR_CUDA if a device-side
 * wait gave up (watchdog): the results of that run are invalid. */
int nwb_plan_summary(nwb_plan *p, nwb_summary *out);
/* Digest of this plan's share of the arrow table on the device (definition: nwb_arrow_digest() above; the
 * shares of a strip group add up mod 2^64).  Blocking. */
int nwb_plan_arrow_digest(nwb_plan *p, uint64_t *digest);
/* "" / "fused" / "dense" / "sparse": how the last run obtained the count (nwb_summary.count_path). */
const char *nwb_plan_count_path_name(const nwb_plan *p);
/* Device pointer / pitch of the arrow table (layout above). */
void *nwb_plan_arrows_device(nwb_plan *p);
size_t nwb_plan_arrow_pitch(const nwb_plan *p);
/* Copy arrow rows [row_begin,row_end) (0-based interior rows, j-1) to host. */
int nwb_plan_download_arrows(nwb_plan *p, uint8_t *dst, size_t dst_pitch, int row_begin, int row_end);
/* Kernel launches issued by the plan since creation. */
int64_t nwb_plan_launches(const nwb_plan *p);
/* Device time (ms, CUDA events on the launching stream) of the last run. */
float nwb_plan_kernel_ms(nwb_plan *p);
/* Name of the fill kernel the last run launched ("nwb_fill_hx_kernel", "nwb_fill_pk_kernel",
 * "nwb_fill_i32_kernel"; "" before the first run): for logs and profiles. */
const char *nwb_plan_kernel_name(const nwb_plan *p);
/* Interior columns [begin,end) (0-based, i-1) this plan's strips cover. */
int nwb_plan_strip_range(const nwb_plan *p, int *begin_col, int *end_col);
/* Host-only helpers for a launcher with one process per GPU (no device needed).
 * nwb_strip_partition: the interior columns [begin,end) rank `rank` of `world` owns when the table is
 * cut into strips of `strip_width` columns (256 for the packed kernels) -- the contiguous analogue of the
 * reference's per-thread column sets (needleman-wunsch.c:568-571); the same rule nwb_plan_create() uses.
 * nwb_batch_partition: the contiguous pair range of a rank when a batch is sharded with no communication.
 * nwb_strip_group_score: the optimal score of a strip group from the sum of the ranks' partial_r. */
int nwb_strip_partition(int top_len, int strip_width, int rank, int world, int *begin_col, int *end_col);
int nwb_batch_partition(int64_t n_pairs, int rank, int world, int64_t *first_pair, int64_t *pair_count);
int32_t nwb_strip_group_score(int64_t partial_r_sum, int top_len, int side_len, int d);
/* Re-arm the inbound boundary flag before the next run of a strip group
 * (all ranks must do this, then synchronise, before any rank runs again). */
int nwb_plan_reset_inbox(nwb_plan *p, void *stream);
/* A strip group working through a QUEUE of fills (one rank per GPU): like nwb_plan_run(), but consecutive runs need
 * neither nwb_plan_reset_inbox() nor a barrier between the ranks, so rank r starts fill e + 1 while the ranks to its
 * right are still on fill e (the group's throughput is bounded by one rank's share of a fill, not by the
 * whole strip-to-strip wavefront; the reference's analogue is its -p workers picking up the next column set,
 * needleman-wunsch.c:557-574, across consecutive needleman_wunsch() calls, :654-689).  The inbox is double-buffered
 * (fill e streams into copy e & 1 of the right neighbour's inbox) and a rank re-uses a copy only after the neighbour
 * has acknowledged, through a word in its HBM, that it is done with it; that wait is watchdog-bounded (NWB_ERR_CUDA
 * from nwb_plan_summary()).  Every rank of the group must make the same sequence of calls; after any nwb_plan_run()
 * on the same plans call nwb_plan_reset_inbox() on every rank and synchronise before the first pipelined run. */
int nwb_plan_run_pipelined(nwb_plan *p, int m, int k, int d, void *stream);
/* Same-process strip group (one host thread driving several GPUs): make `right`'s inbox the target of `p`'s last
 * strip through CUDA peer access (the in-process counterpart of nwb_plan_ipc_attach_right()). */
int nwb_plan_attach_right(nwb_plan *p, nwb_plan *right);
/* Multi-process column strips: export this rank's inbound boundary buffer as
 * an opaque CUDA IPC blob (nwb_plan_ipc_size() bytes) and attach the right
 * neighbour's blob so this rank's last strip streams its boundary column
 * straight into the neighbour's HBM over NVLink. */
size_t nwb_plan_ipc_size(void);
int nwb_plan_ipc_export(nwb_plan *p, void *blob);
int nwb_plan_ipc_attach_right(nwb_plan *p, const void *blob);

/* ========================================================================== *
 * 3. Batch of independent pairs -- the reference looped per pair
 *    (alloc/init_computation + compute_table_scores + free_computation per
 *    pair); one warp per pair, no inter-block sync.
 * ========================================================================== */
typedef struct nwb_batch nwb_batch;

/* Pair p has top = tops + top_off[p] .. top_off[p+1], side likewise (offset
 * arrays hold n_pairs + 1 entries).  Results per pair: optimal score, branch
 * count (unless NWB_NO_BRANCH_COUNT) and, with NWB_WANT_ARROWS_HOST, the pair's
 * arrow table (layout of section 1 with pitch 128 * ceil(A/256) bytes); with
 * NWB_WANT_COUNT the number of optimal alignments mod 2^64 (a second pass over
 * the arrow codes, nwb_batch_count.cuh; nwb_batch_count_u64()).
 * Any m / k / d is accepted, as by the reference (needleman-wunsch.c:783-785): schemes whose per-cell differences
 * are small run the packed 16x2 batch kernels; everything else, NWB_FORCE_GENERAL, NWB_WANT_SCORES (the int32
 * score matrix of every pair: nwb_batch_score_rows()) and NWB_TRACK_ABS (nwb_batch_greatest_abs()) run the general
 * int32 engine with one warp per pair (nwb_batch_i32.cuh).  Only NWB_WANT_COUNT_MATRIX is refused
 * (NWB_ERR_UNSUPPORTED: use nwb_fill() per pair).
 * nwb_fill_batch() = nwb_batch_create() + nwb_batch_run() + nwb_batch_fetch(). */
int nwb_fill_batch(const char *tops, const int64_t *top_off,
                   const char *sides, const int64_t *side_off, int64_t n_pairs,
                   int m, int k, int d, unsigned flags, int device, nwb_batch **out);
/* Upload the pairs and allocate the device-side tables (strings stay in HBM). */
int nwb_batch_create(const char *tops, const int64_t *top_off,
                     const char *sides, const int64_t *side_off, int64_t n_pairs,
                     int m, int k, int d, unsigned flags, int device, nwb_batch **out);
/* Launch the batch on `stream` (NULL = the batch's own); asynchronous. */
int nwb_batch_run(nwb_batch *b, void *stream);
/* New strings for the SAME shapes (the offsets given to nwb_batch_create()), from host buffers, and run: the batch
 * is cut into chunks of pairs and the host-to-device copy of a chunk overlaps the kernels of the previous one.
 * Asynchronous on the batch's own stream; follow with nwb_batch_fetch().  (nwb_fill_batch() uses it too.) */
int nwb_batch_refill(nwb_batch *b, const char *tops, const char *sides);
/* Wait and copy the per-pair results (and arrows if requested) to the host. */
int nwb_batch_fetch(nwb_batch *b);
void nwb_batch_free(nwb_batch *b);
int64_t nwb_batch_size(const nwb_batch *b);
int32_t nwb_batch_opt_score(const nwb_batch *b, int64_t pair);
uint32_t nwb_batch_branch_count(const nwb_batch *b, int64_t pair);
uint64_t nwb_batch_count_u64(const nwb_batch *b, int64_t pair);
const uint8_t *nwb_batch_arrow_rows(const nwb_batch *b, int64_t pair, size_t *pitch);
/* NWB_TRACK_ABS: max |score| over the pair's interior cells (score_table_t.greatest_abs_val under tflag). */
int32_t nwb_batch_greatest_abs(const nwb_batch *b, int64_t pair);
/* NWB_WANT_SCORES: interior scores of a pair, element (i,j), i,j >= 1, at rows[(j-1) * (*pitch_elems) + (i-1)]. */
const int32_t *nwb_batch_score_rows(const nwb_batch *b, int64_t pair, size_t *pitch_elems);
float nwb_batch_kernel_ms(const nwb_batch *b);
int64_t nwb_batch_launches(const nwb_batch *b);
/* Name of the kernel nwb_batch_run() launches for this batch: "nwb_batch_bp_kernel" (bit-parallel rows, one thread
 * per pair: top strings of at most 256 characters and 2d + m <= 3, e.g. DNA 1/1/1; pairs whose top string has more
 * than five distinct letters are worked off by nwb_batch_pk_kernel right behind it), "nwb_batch_bx_kernel" (two
 * pairs per warp: top strings of at most 256 characters and 2d + m <= 7), "nwb_batch_cx_kernel" (the same, pairs
 * swept back to back: every pair has the same shape and the side length is a multiple of 32),
 * "nwb_batch_i32_kernel" (any m / k / d, scores), "nwb_batch_pk_kernel" otherwise.  For logs and profiles. */
const char *nwb_batch_kernel_name(const nwb_batch *b);
void *nwb_batch_arrows_device(nwb_batch *b);
/* Parity aid: digests of the whole batch computed on the device (after nwb_batch_run): out[0..3] = the sums
 * mod 2^64 over the pairs p of mix64(first_pair + p, x_p), x_p = the pair's arrow digest (nwb_arrow_digest) /
 * optimal score (sign-extended) / branch count / alignment count (0 terms when not produced).  first_pair
 * = the global index of this batch's pair 0, so that the shards of a sharded batch add up.  Blocking. */
int nwb_batch_digest(nwb_batch *b, int64_t first_pair, uint64_t out[4]);

/* ========================================================================== *
 * 4. Measurement aid (not on the fill path): INT/DPX issue rate of `device`,
 *    from independent VIADDMNMX (mode 0) / VIMNMX3 (1) / VIMNMX3.U16x2 (2) /
 *    VIMNMX3+IMAD (3) chains at full occupancy.  Returns
 *    thread-instructions per clock per SM (in-kernel clock64) and G/s (events).
 * ========================================================================== */
int nwb_measure_int_issue(int device, int mode, double *per_clk_per_sm, double *gops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* NWB_H */
