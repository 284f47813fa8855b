/*
 * nwb_fill_i32.cuh -- general score-table fill: int32 scores, any m/k/d.
 *
 * Replaces score_cell()/score_cell_column()/score_cell_column_set()/
 * compute_table_scores() (reference needleman-wunsch.c:418-626) and the border
 * initialisation of init_computation_tables() (computation.c:75-125).
 *
 * Mapping.  The table is cut into column strips of 256 interior columns.  One
 * warp owns one strip and sweeps it top to bottom; lane l holds 8 consecutive
 * columns in registers and runs one row behind lane l-1 (anti-diagonal skew),
 * receiving its left neighbour cell through __shfl_up_sync.  Strips are handed
 * out cyclically to the persistent warps (the reference's cyclic column sets,
 * needleman-wunsch.c:568-571, at strip granularity); strip c waits on strip
 * c-1's progress word (ld.acquire) and reads its boundary column, which strip
 * c-1's lane 31 streamed out (plain stores + st.release every 32 rows).
 *
 * Per cell: diag add, one DPX max-of-three (VIMNMX3), three tie tests that set
 * EVERY arrow equal to the maximum (needleman-wunsch.c:485-503), optional
 * 64-bit path-count adds.  Arrow nibbles (8 per lane and row = one 32-bit
 * word) are transposed through a per-warp shared-memory ring and written with
 * coalesced 128-bit stores.
 */
#pragma once
#include "nwb_device.cuh"

#define NWB_I32_K 8                      /* columns per lane           */
#define NWB_I32_STRIP_W (32 * NWB_I32_K) /* 256 columns per strip      */
#define NWB_I32_RING_ROWS 64
#define NWB_I32_STAGE_WORDS (NWB_I32_RING_ROWS * 32)
#define NWB_I32_WARPS 4
#define NWB_I32_SMEM_BYTES (NWB_I32_WARPS * NWB_I32_STAGE_WORDS * 4)

template <bool COUNT, bool SCORES, bool ABS, bool CNTMAT>
__device__ __forceinline__ bool nwb_i32_strip(const NwbStripParams &p, const int c, uint32_t *stage,
                                               const int lane, unsigned &branches, int &gabs)
{
    const int A = p.A, B = p.B;
    const int m = p.m, nk = -p.k, d = p.d;
    const int c0 = c * NWB_I32_STRIP_W + lane * NWB_I32_K + 1; /* first column (1-based) of this lane */

    /* top characters and column validity of this lane */
    int tc[NWB_I32_K];
    unsigned vmask = 0; /* nibble mask of valid columns */
#pragma unroll
    for (int k = 0; k < NWB_I32_K; k++) {
        const int i = c0 + k;
        tc[k] = (i <= A) ? (int)p.top[i - 1] : -1;
        if (i <= A) vmask |= 0xFu << (4 * k);
    }

    const int lc = c - p.strip_begin; /* local strip index */
    const bool has_left = (c > 0);
    const bool left_remote = has_left && (lc == 0);
    const bool publish = (c + 1 < p.n_strips);
    const bool out_remote = publish && (c == p.strip_end - 1);
    int32_t *out_s = out_remote ? p.out_bnd_s : p.bnd_s + (size_t)lc * p.bpitch;
    unsigned long long *out_c = !COUNT ? nullptr : (out_remote ? p.out_bnd_c : p.bnd_c + (size_t)lc * p.bpitch);
    int *out_flag = out_remote ? p.out_progress : p.progress + lc;
    const int32_t *in_s = nullptr;
    const unsigned long long *in_c = nullptr;
    const int *in_flag = nullptr;
    if (has_left) {
        in_s = left_remote ? p.in_bnd_s : p.bnd_s + (size_t)(lc - 1) * p.bpitch;
        if (COUNT) in_c = left_remote ? p.in_bnd_c : p.bnd_c + (size_t)(lc - 1) * p.bpitch;
        in_flag = left_remote ? p.in_progress : p.progress + (lc - 1);
    }

    int S[NWB_I32_K];                 /* scores of my columns in the previous row */
    unsigned long long Cn[NWB_I32_K]; /* counts of my columns in the previous row */
    int sdiag0 = 0;                   /* S[c0-1][j-1] */
    unsigned long long cdiag0 = 1;
#pragma unroll
    for (int k = 0; k < NWB_I32_K; k++) { S[k] = 0; Cn[k] = 1; }

    const int nblocks = (B + 31 + 31) / 32;
    for (int blk = 0; blk < nblocks; blk++) {
        /* boundary values for lane 0's rows 32*blk+1 .. 32*blk+32 */
        int bq_s = 0;
        unsigned long long bq_c = 1;
        if (has_left) {
            int need = 32 * blk + 32;
            if (need > B) need = B;
            if (!NWB_DBG_BITS(p, 1) && !nwb_wait_ge_wd(in_flag, need, left_remote, &p.summary->error, p.watchdog_ns))
                return false; /* watchdog: the left strip never got there */
            const int jj = 32 * blk + 1 + lane;
            if (jj <= B) {
                bq_s = in_s[jj];
                if (COUNT) bq_c = in_c[jj];
            }
        }
#pragma unroll 1
        for (int t = 0; t < 32; t++) {
            const int s = 32 * blk + t;
            const int j = s - lane + 1;
            /* left neighbour cell (c0-1, j) */
            int left0 = __shfl_up_sync(NWB_FULL_MASK, S[NWB_I32_K - 1], 1);
            unsigned long long cleft0 = 1;
            if (COUNT) cleft0 = __shfl_up_sync(NWB_FULL_MASK, Cn[NWB_I32_K - 1], 1);
            if (has_left) {
                const int b_s = __shfl_sync(NWB_FULL_MASK, bq_s, t);
                unsigned long long b_c = 1;
                if (COUNT) b_c = __shfl_sync(NWB_FULL_MASK, bq_c, t);
                if (lane == 0) { left0 = b_s; cleft0 = b_c; }
            } else if (lane == 0) {
                left0 = -j * d; /* column 0: computation.c:118 */
                cleft0 = 1;
            }
            if (j >= 1 && j <= B) {
                if (j == 1) {
                    /* row 0: computation.c:106 */
#pragma unroll
                    for (int k = 0; k < NWB_I32_K; k++) { S[k] = -(c0 + k) * d; Cn[k] = 1; }
                    sdiag0 = -(c0 - 1) * d;
                    cdiag0 = 1;
                }
                const int sc = (int)p.side[j - 1];
                unsigned word = 0;
                int left = left0 - d;
                int dg = sdiag0;
                unsigned long long cl = cleft0, cd = cdiag0;
#pragma unroll
                for (int k = 0; k < NWB_I32_K; k++) {
                    const int up = S[k] - d;
                    const int diag = dg + ((tc[k] == sc) ? m : nk);
                    const int v = __vimax3_s32(diag, up, left);
                    const bool ad = (v == diag), al = (v == left), au = (v == up);
                    word |= (ad ? 1u : 0u) << (4 * k);
                    word |= (al ? 2u : 0u) << (4 * k);
                    word |= (au ? 4u : 0u) << (4 * k);
                    if (COUNT) {
                        unsigned long long cn = ad ? cd : 0ull;
                        if (al) cn += cl;
                        if (au) cn += Cn[k];
                        cd = Cn[k];
                        Cn[k] = cn;
                        cl = cn;
                    }
                    if (ABS) {
                        const int a = v < 0 ? -v : v;
                        if (((vmask >> (4 * k)) & 1u) && a > gabs) gabs = a;
                    }
                    dg = S[k];
                    S[k] = v;
                    left = v - d;
                }
                sdiag0 = left0;
                cdiag0 = cleft0;
                stage[(j & (NWB_I32_RING_ROWS - 1)) * 32 + lane] = word;
                {
                    /* branch = interior cell with >= 2 arrows (needleman-wunsch.c:507-509) */
                    const unsigned w = word & vmask;
                    const unsigned b0 = w & 0x11111111u, b1 = (w >> 1) & 0x11111111u, b2 = (w >> 2) & 0x11111111u;
                    branches += (unsigned)__popc((b0 & b1) | (b0 & b2) | (b1 & b2));
                }
                if (SCORES) {
                    int32_t *row = p.scores + (size_t)(j - 1) * p.spitch + (c0 - 1);
#pragma unroll
                    for (int k = 0; k < NWB_I32_K; k++) row[k] = S[k];
                }
                if (CNTMAT) {
                    unsigned long long *row = p.cntmat + (size_t)(j - 1) * p.spitch + (c0 - 1);
#pragma unroll
                    for (int k = 0; k < NWB_I32_K; k++) row[k] = Cn[k];
                }
                if (lane == 31 && publish) {
                    out_s[j] = S[NWB_I32_K - 1];
                    if (COUNT) out_c[j] = Cn[NWB_I32_K - 1];
                }
                if (j == B && c0 <= A && A < c0 + NWB_I32_K) {
                    /* the lane that owns column A: cells[M-1][N-1] */
                    int fs = 0;
                    unsigned long long fc = 0;
#pragma unroll
                    for (int k = 0; k < NWB_I32_K; k++)
                        if (c0 + k == A) { fs = S[k]; fc = Cn[k]; }
                    p.summary->opt_score = fs;
                    p.summary->count = COUNT ? fc : 0ull;
                }
            }
        }
        __syncwarp();
        /* rows <= 32*blk+1 are complete in the ring: flush the 32 newest complete rows */
        {
            const int jhi = 32 * blk + 1;
            const int jlo = jhi - 31;
            const int quad = lane & 7;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int j = jlo + r * 4 + (lane >> 3);
                if (j >= 1 && j <= B) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(stage + (j & (NWB_I32_RING_ROWS - 1)) * 32 + quad * 4);
                    *reinterpret_cast<uint4 *>(p.arrows + (size_t)(j - 1) * p.pitch + (size_t)c * (NWB_I32_STRIP_W / 2) + quad * 16) = v;
                }
            }
        }
        __syncwarp();
        if (publish && lane == 31 && !NWB_FAULT_INJECTED(p)) {
            int done = 32 * blk + 1;
            if (done > B) done = B;
            if (out_remote) {
                __threadfence_system();
                nwb_st_release_sys(out_flag, done);
            } else {
                nwb_st_release_gpu(out_flag, done);
            }
        }
    }
    return true;
}

template <bool COUNT, bool SCORES, bool ABS, bool CNTMAT>
__global__ void __launch_bounds__(32 * NWB_I32_WARPS, 1) nwb_fill_i32_kernel(const NwbStripParams p)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nworkers = (int)gridDim.x * (int)(blockDim.x >> 5);
    const int worker = warp * (int)gridDim.x + (int)blockIdx.x;
    uint32_t *stage = reinterpret_cast<uint32_t *>(NWB_SMEM_BASE()) + warp * NWB_I32_STAGE_WORDS;

    unsigned branches = 0;
    int gabs = 0;
    for (int c = p.strip_begin + worker; c < p.strip_end; c += nworkers)
        if (!nwb_i32_strip<COUNT, SCORES, ABS, CNTMAT>(p, c, stage, lane, branches, gabs)) return; /* watchdog */

    /* warp-reduce the per-lane counters, one atomic per warp (replaces the
     * rwlock-guarded inc_branch_count(), walk-table.c:108-120) */
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        branches += __shfl_xor_sync(NWB_FULL_MASK, branches, o);
        const int g2 = __shfl_xor_sync(NWB_FULL_MASK, gabs, o);
        if (g2 > gabs) gabs = g2;
    }
    if (lane == 0) {
        if (branches) atomicAdd(&p.summary->branch_count, branches);
        if (ABS && gabs) atomicMax(&p.summary->greatest_abs, gabs);
    }
}
