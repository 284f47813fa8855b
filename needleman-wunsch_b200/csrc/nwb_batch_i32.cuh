/*
 * nwb_batch_i32.cuh -- batch of independent pairs through the GENERAL int32 engine.
 *
 * The reference takes any m / k / d that atoi() yields (needleman-wunsch.c:783-785);
 * the packed batch kernels (nwb_batch.cuh, nwb_batch_bx.cuh) cover the schemes whose
 * per-cell differences are small.  Everything else -- negative or large penalties,
 * NWB_FORCE_GENERAL, the int32 score matrix behind `-t` (NWB_WANT_SCORES) and the
 * interior |score| maximum (NWB_TRACK_ABS) -- runs here: one warp per pair, the
 * pair's 256-column strips swept one after the other by nwb_i32_strip() (the same
 * device code as nwb_fill_i32_kernel: score_cell(), needleman-wunsch.c:418-510),
 * the boundary column between two strips kept in a per-warp scratch line.
 */
#pragma once
#include "nwb_fill_i32.cuh"

#define NWB_BI32_WARPS 8

struct NwbBatchI32Params {
    const uint8_t *tops;
    const long long *top_off;
    const uint8_t *sides;
    const long long *side_off;
    long long n_pairs;
    int m, k, d;
    uint8_t *arrows;            /* per pair: B rows x (128 * ceil(A/256)) bytes at arrow_off[p] */
    const long long *arrow_off;
    int32_t *scores;            /* NWB_WANT_SCORES: per pair B rows x (256 * ceil(A/256)) int32 at score_off[p] */
    const long long *score_off;
    int *out_score;
    unsigned *out_branch;       /* or NULL */
    int *out_abs;               /* NWB_TRACK_ABS, or NULL */
    int32_t *bnd_s;             /* per warp: max_strips * bpitch boundary scores */
    size_t bpitch;
    int *progress;              /* per warp: max_strips progress words */
    int max_strips;
    NwbDevSummary *wsum;        /* per warp */
};

template <bool SCORES, bool ABS>
__global__ void __launch_bounds__(32 * NWB_BI32_WARPS) nwb_batch_i32_kernel(const NwbBatchI32Params bp)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * NWB_BI32_WARPS + warp;
    const long long nwarps = (long long)gridDim.x * NWB_BI32_WARPS;
    uint32_t *stage = reinterpret_cast<uint32_t *>(NWB_SMEM_BASE()) + warp * NWB_I32_STAGE_WORDS;
    NwbDevSummary *wsum = bp.wsum + gw;
    int *progress = bp.progress + gw * bp.max_strips;

    for (long long pr = gw; pr < bp.n_pairs; pr += nwarps) {
        const long long t0 = bp.top_off[pr], s0 = bp.side_off[pr];
        const int A = (int)(bp.top_off[pr + 1] - t0), B = (int)(bp.side_off[pr + 1] - s0);
        if (A == 0 || B == 0) {
            /* borders only (computation.c:97-124) */
            if (lane == 0) {
                bp.out_score[pr] = (A == 0) ? -B * bp.d : -A * bp.d;
                if (bp.out_branch) bp.out_branch[pr] = 0u;
                if (bp.out_abs) bp.out_abs[pr] = 0;
            }
            continue;
        }
        const int n_strips = (A + NWB_I32_STRIP_W - 1) / NWB_I32_STRIP_W;
        for (int c = lane; c < n_strips; c += 32) progress[c] = 0;
        __syncwarp();

        NwbStripParams sp;
        sp.top = bp.tops + t0;
        sp.side = bp.sides + s0;
        sp.side_pre = nullptr;
        sp.A = A; sp.B = B; sp.m = bp.m; sp.k = bp.k; sp.d = bp.d;
        sp.n_strips = n_strips;
        sp.strip_begin = 0;
        sp.strip_end = n_strips;
        sp.arrows = bp.arrows + bp.arrow_off[pr];
        sp.pitch = (size_t)n_strips * (NWB_I32_STRIP_W / 2);
        sp.scores = SCORES ? bp.scores + bp.score_off[pr] : nullptr;
        sp.cntmat = nullptr;
        sp.spitch = (size_t)n_strips * NWB_I32_STRIP_W;
        sp.bnd_s = bp.bnd_s + (size_t)gw * bp.max_strips * bp.bpitch;
        sp.bnd_c = nullptr; sp.bnd_w = nullptr;
        sp.bpitch = bp.bpitch;
        sp.progress = progress;
        sp.in_bnd_s = nullptr; sp.in_bnd_c = nullptr; sp.in_bnd_w = nullptr; sp.in_progress = nullptr;
        sp.out_bnd_s = nullptr; sp.out_bnd_c = nullptr; sp.out_bnd_w = nullptr; sp.out_progress = nullptr;
        sp.summary = wsum;
        sp.count_branches = 0;
        sp.publish_rows = 0;
        sp.debug_nowait = 0;
        sp.watchdog_ns = ~0ull; /* the left strip was finished by this very warp: its words are always there */
        sp.debug_times = nullptr; sp.debug_trace = nullptr; sp.debug_trace_stride = 1; sp.debug_trace_blocks = 0;

        unsigned branches = 0;
        int gabs = 0;
        for (int c = 0; c < n_strips; c++) {
            nwb_i32_strip<false, SCORES, ABS, false>(sp, c, stage, lane, branches, gabs);
            __syncwarp();
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            branches += __shfl_xor_sync(NWB_FULL_MASK, branches, o);
            const int g2 = __shfl_xor_sync(NWB_FULL_MASK, gabs, o);
            if (g2 > gabs) gabs = g2;
        }
#ifndef NWB_EMU
        __threadfence_block(); /* the lane that owns column A wrote wsum->opt_score */
#endif
        __syncwarp();
        if (lane == 0) {
            bp.out_score[pr] = *reinterpret_cast<volatile int *>(&wsum->opt_score);
            if (bp.out_branch) bp.out_branch[pr] = branches;
            if (bp.out_abs) bp.out_abs[pr] = gabs;
        }
        __syncwarp();
    }
}
