/*
 * nwb_batch.cuh -- batch of independent pairs: one warp per pair, no inter-block
 * synchronisation (BASELINE config 4: 1M pairs of 256 x 256).  Replaces the
 * reference looped per pair: alloc_computation + init_computation +
 * compute_table_scores + free_computation (computation.c:51-214,
 * needleman-wunsch.c:583) for every pair.
 *
 * A warp runs the packed 16x2 strip engine (nwb_fill_pk.cuh, K = 4 columns per
 * half-lane, one row per step) over the 256-column strips of its pair, left to
 * right; a pair of up to 256 columns is a single strip and needs no boundary
 * stream at all.  The pair's side string is pre-shifted into the warp's shared
 * memory once.  Pairs are handed out grid-stride.
 */
#pragma once
#include "nwb_fill_pk.cuh"

#define NWB_BATCH_WARPS 12
#define NWB_BATCH_K 4
#define NWB_BATCH_R 1
#define NWB_BATCH_SPADB 64   /* side_pre entries in front of row 1 (lane 31 starts 61 rows above the table) */
#define NWB_BATCH_STAIL 192  /* ... and behind row B (skew + block rounding + prefetch) */

struct NwbBatchParams {
    const uint8_t *tops;
    const long long *top_off;   /* n_pairs + 1 */
    const uint8_t *sides;
    const long long *side_off;  /* n_pairs + 1 */
    long long n_pairs;
    int m, k, d;
    int max_B;                  /* longest side string: sizes the shared side_pre copy */
    uint8_t *arrows;            /* all pairs' nibble tables */
    const long long *arrow_off; /* byte offset of pair p's table; its pitch is 128 * ceil(A_p / 256) */
    int *out_score;             /* [n_pairs] */
    unsigned *out_branch;       /* [n_pairs] branch counters, or NULL */
    uint32_t *scratch;          /* boundary streams for pairs wider than one strip: per warp n_strips_max * bpitch */
    size_t scratch_per_warp;    /* words */
    size_t bpitch;              /* words per strip boundary */
    const long long *pair_list; /* NULL, or the pairs to fill (what nwb_batch_bp_kernel left over) ... */
    const unsigned *pair_count; /* ... and how many of them (device word) */
};

#define NWB_BATCH_SIDE_ELEMS(maxB) ((size_t)(maxB) + NWB_BATCH_SPADB + NWB_BATCH_STAIL)
#define NWB_BATCH_SMEM_PER_WARP(maxB) \
    (NWB_PK_WARP_SMEM(NWB_BATCH_K, NWB_BATCH_R, false) + ((NWB_BATCH_SIDE_ELEMS(maxB) * 2 + 15) / 16) * 16)

__global__ void __launch_bounds__(32 * NWB_BATCH_WARPS, 1) nwb_batch_pk_kernel(const NwbBatchParams bp, const NwbPkConsts pc)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long gwarp = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    const size_t per_warp = NWB_BATCH_SMEM_PER_WARP(bp.max_B);
    unsigned char *stage = NWB_SMEM_BASE() + (size_t)warp * per_warp;
    uint16_t *side_sm = reinterpret_cast<uint16_t *>(stage + NWB_PK_WARP_SMEM(NWB_BATCH_K, NWB_BATCH_R, false));

    const long long n_todo = bp.pair_list ? (long long)*bp.pair_count : bp.n_pairs;
    for (long long q = gwarp; q < n_todo; q += nwarps) {
        const long long pr = bp.pair_list ? bp.pair_list[q] : q;
        const long long t0 = bp.top_off[pr], s0 = bp.side_off[pr];
        const int A = (int)(bp.top_off[pr + 1] - t0), B = (int)(bp.side_off[pr + 1] - s0);
        if (A == 0 || B == 0) {
            if (lane == 0) {
                bp.out_score[pr] = (A == 0) ? -B * bp.d : -A * bp.d; /* borders only */
                if (bp.out_branch) bp.out_branch[pr] = 0u;
            }
            continue;
        }
        /* side_pre of this pair: entry e is row e - SPADB */
        const int nelem = B + NWB_BATCH_SPADB + NWB_BATCH_STAIL;
        for (int e = lane; e < nelem; e += 32) {
            const int j = e - NWB_BATCH_SPADB;
            unsigned v = 0xFFFFu;
            if (j >= 1 && j <= B) v = (~((unsigned)bp.sides[s0 + j - 1] << pc.shift)) & 0xFFFFu;
            side_sm[e] = (uint16_t)v;
        }
        const int n_strips = (A + 64 * NWB_BATCH_K - 1) / (64 * NWB_BATCH_K);
        uint32_t *scr = bp.scratch + (size_t)gwarp * bp.scratch_per_warp;
        if (n_strips > 1) { /* boundary streams start invalid */
            const size_t nw = (size_t)(n_strips - 1) * bp.bpitch;
            for (size_t e = lane; e < nw; e += 32) scr[e] = 0u;
        }
        __syncwarp();

        NwbStripParams sp;
        sp.top = bp.tops + t0;
        sp.side = bp.sides + s0;
        sp.side_pre = side_sm - (NWB_PK_SPAD - NWB_BATCH_SPADB); /* the strip engine indexes from row -NWB_PK_SPAD */
        sp.A = A; sp.B = B; sp.m = bp.m; sp.k = bp.k; sp.d = bp.d;
        sp.n_strips = n_strips;
        sp.strip_begin = 0;
        sp.strip_end = n_strips;
        sp.arrows = bp.arrows + bp.arrow_off[pr];
        sp.pitch = (size_t)n_strips * 32 * NWB_BATCH_K;
        sp.scores = nullptr; sp.cntmat = nullptr; sp.spitch = 0;
        sp.bnd_s = nullptr; sp.bnd_c = nullptr;
        sp.bnd_w = scr;
        sp.bpitch = bp.bpitch;
        sp.progress = nullptr;
        sp.in_bnd_s = nullptr; sp.in_bnd_c = nullptr; sp.in_bnd_w = nullptr; sp.in_progress = nullptr;
        sp.out_bnd_s = nullptr; sp.out_bnd_c = nullptr; sp.out_bnd_w = nullptr; sp.out_progress = nullptr;
        sp.summary = nullptr;
        sp.count_branches = 0;
        sp.debug_nowait = 0;
        sp.debug_times = nullptr; sp.debug_trace = nullptr; sp.debug_trace_stride = 1; sp.debug_trace_blocks = 0;

        long long rsum = 0;
        unsigned long long cfinal_unused = 0ull;
        unsigned branches = 0;
        for (int c = 0; c < n_strips; c++) {
            nwb_pk_strip<NWB_BATCH_K, NWB_BATCH_R, true, false>(sp, pc, c, stage, lane, rsum, nullptr, cfinal_unused,
                                                                bp.out_branch != nullptr, branches);
            __syncwarp();
        }
        if (n_strips > 1) {
            /* the share of r(A,B) left of the last strip: sum of v down the boundary it consumed */
            const uint32_t *w = scr + (size_t)(n_strips - 2) * bp.bpitch + NWB_PK_BPAD;
            for (int g = lane; g < B; g += 32) rsum += (long long)((nwb_ld_relaxed_u32(w + g, false) >> 16) & 0x7FFFu);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rsum += __shfl_xor_sync(NWB_FULL_MASK, rsum, o);
            branches += __shfl_xor_sync(NWB_FULL_MASK, branches, o);
        }
        if (lane == 0 && bp.out_branch) bp.out_branch[pr] = branches;
        /* score(A,B) = sum_i u(i,B) - d*(A+B) */
        if (lane == 0) bp.out_score[pr] = (int)(unsigned)((unsigned long long)rsum - (unsigned long long)((long long)bp.d * ((long long)A + B)));
        __syncwarp();
    }
}
