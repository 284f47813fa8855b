/*
 * nwb_batch_count.cuh -- the optimal-alignment count behind `-s` for a batch of pairs.
 *
 * Per pair the number of arrow paths from (A,B) to (0,0), mod 2^64:
 *     cnt(0,0) = cnt(i,0) = cnt(0,j) = 1,
 *     cnt(i,j) = [DIAG] cnt(i-1,j-1) + [LEFT] cnt(i-1,j) + [UP] cnt(i,j-1)
 * (the reference enumerates every alignment, needleman-wunsch.c:209-331, and keeps the low 32 bits,
 * computation.c:223-260).  Same recurrence and row kernel as nwb_count.cuh (nwb_count_row<8>: six selects and
 * one three-input add with carry per cell), here as a second pass over the 4-bit codes a batch fill kernel
 * has written: one warp per pair, lane l owns columns 8l+1 .. 8l+8 of a 256-column strip and works on row
 * t - l + 1 at step t; the strips of a wider pair are swept left to right by the same warp, the last column's
 * counts of a strip waiting in a per-warp scratch line for the next one.  Arrow rows come in with coalesced
 * 16-byte loads (four whole 128-byte strip rows per instruction) four steps ahead of lane 0 and are handed
 * to the skewed lanes through a 64-row shared-memory ring.  No inter-warp synchronisation.
 */
#pragma once
#include "nwb_count.cuh"
#include "nwb_batch.cuh"

#define NWB_BCNT_WARPS 16
#define NWB_BCNT_RING_ROWS 64
#define NWB_BCNT_SMEM_PER_WARP (NWB_BCNT_RING_ROWS * 128)

struct NwbBatchCountParams {
    const long long *top_off;   /* n_pairs + 1 */
    const long long *side_off;  /* n_pairs + 1 */
    long long n_pairs;
    const uint8_t *arrows;      /* all pairs' nibble tables */
    const long long *arrow_off; /* byte offset of pair p's table; its pitch is 128 * ceil(A_p / 256) */
    unsigned long long *out_count; /* [n_pairs] */
    unsigned long long *scratch;   /* pairs wider than one strip: per warp max_B + 1 boundary counts */
    size_t scratch_per_warp;       /* elements */
    const long long *pair_list;    /* NULL, or the pairs to count (what nwb_batch_lcount_kernel left over) ... */
    const unsigned *pair_count;    /* ... and how many of them (device word) */
};

__global__ void __launch_bounds__(32 * NWB_BCNT_WARPS, 1) nwb_batch_count_kernel(const NwbBatchCountParams cp)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long gwarp = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    unsigned *ring = reinterpret_cast<unsigned *>(NWB_SMEM_BASE() + (size_t)warp * NWB_BCNT_SMEM_PER_WARP);
    unsigned long long *bnd = cp.scratch ? cp.scratch + (size_t)gwarp * cp.scratch_per_warp : nullptr;
    const int sub = lane >> 3, chunk = lane & 7;

    const long long n_todo = cp.pair_list ? (long long)*cp.pair_count : cp.n_pairs;
    for (long long q = gwarp; q < n_todo; q += nwarps) {
        const long long pr = cp.pair_list ? cp.pair_list[q] : q;
        const int A = (int)(cp.top_off[pr + 1] - cp.top_off[pr]), B = (int)(cp.side_off[pr + 1] - cp.side_off[pr]);
        if (A == 0 || B == 0) {
            if (lane == 0) cp.out_count[pr] = 1ull; /* borders only: one path along the border */
            continue;
        }
        const int n_strips = (A + 255) / 256;
        const size_t pitch = (size_t)n_strips * 128;
        const uint8_t *tab = cp.arrows + cp.arrow_off[pr];
        for (int c = 0; c < n_strips; c++) {
            const uint8_t *src = tab + (size_t)c * 128 + (size_t)chunk * 16;
            /* the cell (A,B): lane / column that owns it, in the last strip */
            const int kfin = (c == n_strips - 1) ? (A - 1 - c * 256) - 8 * lane : -1;
            unsigned long long cnt[8];
#pragma unroll
            for (int k = 0; k < 8; k++) cnt[k] = 1ull; /* border row */
            unsigned long long send = 1ull, left_above = 1ull;
            __syncwarp(); /* the previous strip's / pair's ring reads and boundary writes are done */
            /* rows 1..4 into the ring (slot = row mod 64) */
            {
                const int r = 1 + sub;
                uint4 w = make_uint4(0u, 0u, 0u, 0u);
                if (r <= B) w = nwb_ldg_u128(src + (size_t)(r - 1) * pitch);
                *reinterpret_cast<uint4 *>(ring + (r & (NWB_BCNT_RING_ROWS - 1)) * 32 + chunk * 4) = w;
            }
            uint4 wnext = nwb_ldg_u128(src + (size_t)((5 + sub <= B ? 5 + sub : B) - 1) * pitch); /* rows 5..8, stored before step 4 */
            __syncwarp();
            const int nsteps = B + 31;
            const bool has_left = (c > 0), publish = (lane == 31) && (c + 1 < n_strips);
            /* rows t+1 .. t+4 (loaded four steps ago) into the ring; rows t+5 .. t+8 on their way.  The slots they
             * overwrite held rows t-63 .. t-60, last read by lane 31 at step t-30 at the latest. */
            auto stage = [&](const int t) {
                const int r = t + 1 + sub;
                *reinterpret_cast<uint4 *>(ring + (r & (NWB_BCNT_RING_ROWS - 1)) * 32 + chunk * 4) = wnext;
                const int r2 = r + 4;
                /* unconditional (rows below the table re-read row B and are never looked at): a conditional load
                 * goes through a temporary whose copy waits for the load right here */
                wnext = nwb_ldg_u128(src + (size_t)((r2 <= B ? r2 : B) - 1) * pitch);
                __syncwarp();
            };
            /* checked step: my row may lie above row 1 or below row B */
            auto step_checked = [&](const int t) {
                const int j = t - lane + 1;
                unsigned long long cl = __shfl_up_sync(NWB_FULL_MASK, send, 1);
                if (lane == 0) cl = (!has_left || j > B) ? 1ull : bnd[j]; /* column 0 of the table / the strip to my left */
                if (j >= 1 && j <= B) {
                    const unsigned x = ring[(j & (NWB_BCNT_RING_ROWS - 1)) * 32 + lane];
                    nwb_count_row<8>(x, cnt, left_above, cl, send);
                    if (j == B && kfin >= 0 && kfin < 8) {
#pragma unroll
                        for (int k = 0; k < 8; k++)
                            if (k == kfin) cp.out_count[pr] = cnt[k];
                    }
                    if (publish) bnd[j] = send; /* read by lane 0 in the next strip */
                }
            };
            int t = 0;
#pragma unroll 1
            for (; t < 32 && t < nsteps; t++) { /* head: lanes l > t are still above row 1 */
                if ((t & 3) == 0 && t > 0) stage(t);
                step_checked(t);
            }
            /* steps 32 .. B-2: every lane strictly inside rows 1 .. B-1, four steps per staged block of rows */
            const unsigned *rq = ring + lane;
#pragma unroll 1
            for (; t + 4 <= B - 1; t += 4) {
                stage(t);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int j = t + i - lane + 1;
                    unsigned long long cl = __shfl_up_sync(NWB_FULL_MASK, send, 1);
                    if (lane == 0) cl = has_left ? bnd[j] : 1ull;
                    const unsigned x = rq[(j & (NWB_BCNT_RING_ROWS - 1)) * 32];
                    nwb_count_row<8>(x, cnt, left_above, cl, send);
                    if (publish) bnd[j] = send;
                }
            }
#pragma unroll 1
            for (; t < nsteps; t++) { /* tail: the last rows, lanes running out of the table */
                if ((t & 3) == 0) stage(t);
                step_checked(t);
            }
        }
    }
}

/*
 * Uniform batches (every pair A x B with A <= 256, B >= 64, B a multiple of 4 -- BASELINE config 4): a warp takes a
 * contiguous run of pairs and sweeps their arrow tables BACK TO BACK as one tall table of n * B rows (the tables of
 * consecutive pairs are contiguous in memory: one strip, pitch 128).  Row 1 of a table needs nothing from the table
 * above but the border values, so lane l simply resets its counts when it enters a new table (tall row = 1 mod B) and
 * stores the finished pair's count when it leaves one (tall row = 0 mod B): the 31 steps of lane skew are paid once per
 * warp instead of once per pair (B + 31 -> B steps per pair), the ring is primed once, and only the chunks of four
 * steps in which some lane crosses a table boundary (9 of B / 4) run the checked body.
 */
static inline bool nwb_bcount_chain_usable(bool uniform, long long A, long long B, long long n_pairs, long long nwarps)
{
    return uniform && A >= 1 && A <= 256 && B >= 64 && B % 4 == 0 && nwarps > 0 && (n_pairs / nwarps + 1) * B < (1LL << 30);
}

__global__ void __launch_bounds__(32 * NWB_BCNT_WARPS, 1) nwb_batch_count_chain_kernel(const NwbBatchCountParams cp, const int A, const int B)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long gwarp = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    unsigned *ring = reinterpret_cast<unsigned *>(NWB_SMEM_BASE() + (size_t)warp * NWB_BCNT_SMEM_PER_WARP);
    const int sub = lane >> 3, chunk = lane & 7;
    /* my run of pairs: the first n_pairs % nwarps warps take one more */
    const long long per = cp.n_pairs / nwarps, rem = cp.n_pairs % nwarps;
    const long long first = gwarp * per + (gwarp < rem ? gwarp : rem);
    const long long npairs = per + (gwarp < rem ? 1 : 0);
    if (npairs == 0) return;
    const int T = (int)(npairs * B); /* rows of the tall table (the host checks that it fits an int) */
    const uint8_t *src = cp.arrows + cp.arrow_off[first] + (size_t)chunk * 16;
    const int kfin = (A - 1) - 8 * lane; /* the column of cell (A, B), 0..7 in the lane that owns it */
    unsigned long long *out = cp.out_count + first;

    unsigned long long cnt[8];
#pragma unroll
    for (int k = 0; k < 8; k++) cnt[k] = 1ull;
    unsigned long long send = 1ull, left_above = 1ull;
    /* rows 1..4 into the ring (slot = tall row mod 64), rows 5..8 on their way */
    {
        const int r = 1 + sub;
        *reinterpret_cast<uint4 *>(ring + (r & (NWB_BCNT_RING_ROWS - 1)) * 32 + chunk * 4) = nwb_ldg_u128(src + (size_t)(r - 1) * 128);
    }
    uint4 wnext = nwb_ldg_u128(src + (size_t)((5 + sub <= T ? 5 + sub : T) - 1) * 128);
    __syncwarp();
    auto stage = [&](const int t) { /* rows t+1 .. t+4 (loaded four steps ago) into the ring; t+5 .. t+8 on their way */
        const int r = t + 1 + sub;
        *reinterpret_cast<uint4 *>(ring + (r & (NWB_BCNT_RING_ROWS - 1)) * 32 + chunk * 4) = wnext;
        const int r2 = r + 4;
        wnext = nwb_ldg_u128(src + (size_t)((r2 <= T ? r2 : T) - 1) * 128);
        __syncwarp();
    };
    int jrow = -lane;        /* my row inside its table, 0-based, at the next step (negative: above the first table) */
    int pair = 0;            /* index of the table my row is in */
    const unsigned *rq = ring + lane;
    const int nsteps = T + 31;
    int c = 0;               /* t mod B */
    for (int t = 0; t < nsteps; t += 4, c = (c + 4 == B) ? 0 : c + 4) {
        if (t > 0) stage(t);
        if (c < 32 || c == B - 4) {
            /* some lane enters or leaves a table in these four steps (or is above the first / below the last one) */
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int tau = t + i - lane + 1; /* my tall row */
                unsigned long long cl = __shfl_up_sync(NWB_FULL_MASK, send, 1);
                if (lane == 0) cl = 1ull; /* column 0 */
                if (tau >= 1 && tau <= T) {
                    if (jrow == 0) { /* row 1 of a table: the border row above it */
#pragma unroll
                        for (int k = 0; k < 8; k++) cnt[k] = 1ull;
                        left_above = 1ull;
                    }
                    const unsigned x = rq[(tau & (NWB_BCNT_RING_ROWS - 1)) * 32];
                    nwb_count_row<8>(x, cnt, left_above, cl, send);
                    if (jrow == B - 1) {
                        if (kfin >= 0 && kfin < 8) {
#pragma unroll
                            for (int k = 0; k < 8; k++)
                                if (k == kfin) out[pair] = cnt[k];
                        }
                        jrow = -1;
                        pair++;
                    }
                }
                jrow++;
            }
        } else {
            /* every lane strictly inside rows 2 .. B-1 of its table */
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int tau = t + i - lane + 1;
                unsigned long long cl = __shfl_up_sync(NWB_FULL_MASK, send, 1);
                if (lane == 0) cl = 1ull;
                const unsigned x = rq[(tau & (NWB_BCNT_RING_ROWS - 1)) * 32];
                nwb_count_row<8>(x, cnt, left_above, cl, send);
            }
            jrow += 4;
        }
    }
}
