/*
 * nwb_batch_lcount.cuh -- the alignment count behind `-s` for a batch, one THREAD per pair: a sparse backward
 * sweep over the arrow codes the fill has written ("lane count").
 *
 * get_solution_count() (computation.c:249-260; the reference enumerates every alignment,
 * needleman-wunsch.c:209-331) is the number of arrow paths from (A,B) to (0,0).  As in nwb_count_sparse.cuh:
 *     P(A,B) = 1,
 *     P(i,j) = [DIAG(i+1,j+1)] P(i+1,j+1) + [LEFT(i+1,j)] P(i+1,j) + [UP(i,j+1)] P(i,j+1)   (mod 2^64),
 * only cells on optimal paths are non-zero, and the count is the flow that reaches the border (row 0 and
 * column 0 carry one forced arrow each, computation.c:97-124).  The dense batch pass (nwb_batch_count.cuh) visits
 * all A*B cells with a warp per pair, 8 instructions per cell; on BASELINE config 4 the live band is 12 columns
 * wide on average (max 45 in 400 pairs).
 *
 * Here a lane keeps a window of NWB_LC_W = 40 columns of the current row in registers (80 registers of counts).
 * Window cell k of row j is column base_j + k, and base_{j-1} = base_j - 1 by construction: DIAG keeps k, UP goes
 * to k + 1, LEFT runs go right to left inside the row, so the update is static register-to-register data flow
 *     inc'[k] = [DIAG(k)] P[k] + [UP(k-1)] P[k-1],    P[k] = inc[k] + [LEFT(k+1)] P[k+1].
 * Every fourth row the lane looks at its guard zones (the outer 8 cells on either side): live cells there move the
 * window by 8 columns for the next row (the arrow words of the next row are requested before the update of this
 * one, so the decision is made one row ahead).  The arrow nibbles of the window are cut out of three aligned
 * 16-byte loads per row.  Column 0 is given the border's forced UP arrow, so flow that has reached it keeps
 * moving up inside the window and is collected with everything else when the sweep arrives in row 0.
 *
 * If the band does not fit (both guard zones live, a LEFT run that leaves the window, ...), the lane gives up:
 * the pair is put on a list and the dense kernel computes it afterwards -- never a wrong count.
 * tools/lcount_proto.py is the executable statement of the sweep (5 % of config 4's pairs give up).
 */
#pragma once
#include "nwb_batch_count.cuh"

#define NWB_LC_W 40      /* window columns (5 nibble words) */
#define NWB_LC_GUARD 8
#define NWB_LC_SHIFT 8
#define NWB_LC_START 28  /* window position of column A in row B */
#define NWB_LC_PERIOD 4  /* rows between two looks at the guard zones */
#define NWB_LC_AHEAD 6   /* rows between the L2 prefetch and the load */
#define NWB_LC_WARPS 12 /* at most, per block (160 registers per thread) */

/* Warps per block: the groups of 32 pairs are worked off in rounds of grid x warps and a round takes about as long
 * with 9 warps per SM as with 12 (the sweep is bound by the ALU pipe and by its own dependent chains): the smallest
 * block that needs the fewest rounds. */
static inline int nwb_lc_choose_warps(long long groups, int grid)
{
    if (groups <= 0 || grid <= 0) return NWB_LC_WARPS;
    const long long rounds = (groups + (long long)grid * NWB_LC_WARPS - 1) / ((long long)grid * NWB_LC_WARPS);
    int w = (int)((groups + rounds * grid - 1) / (rounds * grid));
    if (w < 4) w = 4;
    if (w > NWB_LC_WARPS) w = NWB_LC_WARPS;
    return w;
}

struct NwbLaneCountParams {
    const long long *top_off;   /* n_pairs + 1 */
    const long long *side_off;  /* n_pairs + 1 */
    long long n_pairs;
    const uint8_t *arrows;      /* all pairs' nibble tables */
    const long long *arrow_off; /* byte offset of pair p's table; its pitch is 128 * ceil(A_p / 256) */
    unsigned long long *out_count; /* [n_pairs] */
    long long *fb_list;         /* pairs left to the dense kernel */
    unsigned *fb_count;
};

/* the 16-byte group g (nibble words 4g .. 4g+3) of a row; outside the row: zeros, except that nibble -1 is the
 * border column with its forced UP arrow */
__device__ __forceinline__ uint4 nwb_lc_group(const uint8_t *row, const int g, const int ngroups)
{
    if (g >= 0 && g < ngroups) return *reinterpret_cast<const uint4 *>(row + (size_t)g * 16);
    return make_uint4(0u, 0u, 0u, (g == -1) ? 0x40000000u : 0u);
}

/* pull a 128-byte line towards the L2 (no register, no scoreboard) */
__device__ __forceinline__ void nwb_lc_prefetch(const uint8_t *p)
{
#ifndef NWB_EMU
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

__device__ __forceinline__ unsigned long long nwb_lc_sel(const unsigned flag, const unsigned long long v) { return flag ? v : 0ull; }

__global__ void __launch_bounds__(32 * NWB_LC_WARPS, 1) nwb_batch_lcount_kernel(const NwbLaneCountParams cp)
{
    constexpr int W = NWB_LC_W, NX = NWB_LC_W / 8;
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long gwarp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long groups = (cp.n_pairs + 31) / 32;
    for (long long grp = gwarp; grp < groups; grp += nwarps) {
        const long long p = grp * 32 + lane;
        const bool valid = p < cp.n_pairs;
        int A = 0, B = 0;
        const uint8_t *tab = cp.arrows;
        if (valid) {
            A = (int)(cp.top_off[p + 1] - cp.top_off[p]);
            B = (int)(cp.side_off[p + 1] - cp.side_off[p]);
            tab = cp.arrows + cp.arrow_off[p];
        }
        const int ns = (A + 255) / 256 > 0 ? (A + 255) / 256 : 1;
        const size_t pitch = (size_t)ns * 128;
        const int ngroups = ns * 8;
        /* an empty string: one alignment (all gaps) */
        bool active = valid && A > 0 && B > 0;
        bool bailed = false;
        unsigned long long inc[W];
#pragma unroll
        for (int k = 0; k < W; k++) inc[k] = (k == NWB_LC_START) ? 1ull : 0ull;
        int base = A - NWB_LC_START; /* column of window cell 0 in the current row */
        int j = B;                   /* current row */
        int maxB = active ? B : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int x = __shfl_xor_sync(NWB_FULL_MASK, maxB, o);
            maxB = x > maxB ? x : maxB;
        }
        /* arrow words of the current row: three aligned groups from the one that holds nibble base - 1 */
        uint4 c0, c1, c2;
        {
            const int g0 = (base - 1) >> 5;
            const uint8_t *row = tab + (size_t)(j > 0 ? j - 1 : 0) * pitch;
            c0 = active ? nwb_lc_group(row, g0, ngroups) : make_uint4(0u, 0u, 0u, 0u);
            c1 = active ? nwb_lc_group(row, g0 + 1, ngroups) : make_uint4(0u, 0u, 0u, 0u);
            c2 = active ? nwb_lc_group(row, g0 + 2, ngroups) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll 1
        for (int it = 0; it < maxB; it++) {
            const bool on = active && !bailed && j >= 1;
            if (!on) continue; /* done (row 0 reached: inc holds the result) or given up; the others go on */
            /* --- where the window goes for the next row --- */
            int sh = 0;
            if ((it & (NWB_LC_PERIOD - 1)) == 0) {
                unsigned long long l = 0ull, r = 0ull;
#pragma unroll
                for (int k = 0; k < NWB_LC_GUARD; k++) { l |= inc[k]; r |= inc[W - 1 - k]; }
                if (l != 0ull && r != 0ull) bailed = true;
                sh = (l != 0ull) ? -NWB_LC_SHIFT : ((r != 0ull) ? NWB_LC_SHIFT : 0);
            }
            const int nbase = base - 1 + sh;
            /* --- request the next row's words --- */
            uint4 n0, n1, n2;
            {
                const int g0 = (nbase - 1) >> 5;
                const bool ld = j >= 2;
                const uint8_t *row = tab + (size_t)(ld ? j - 2 : 0) * pitch;
                n0 = ld ? nwb_lc_group(row, g0, ngroups) : make_uint4(0u, 0u, 0u, 0u);
                n1 = ld ? nwb_lc_group(row, g0 + 1, ngroups) : make_uint4(0u, 0u, 0u, 0u);
                n2 = ld ? nwb_lc_group(row, g0 + 2, ngroups) : make_uint4(0u, 0u, 0u, 0u);
            }
            /* ... and start the rows further up on their way from DRAM: the window drifts by one column per row, so
             * the line that holds its middle NWB_LC_AHEAD rows up is the right one nearly always (a 16-byte load one
             * row ahead alone leaves the warp waiting for DRAM a third of the time) */
            if (j > NWB_LC_AHEAD + 1) {
                int gm = (nbase - NWB_LC_AHEAD + W / 2 - 1) >> 5;
                gm = gm < 0 ? 0 : (gm >= ngroups ? ngroups - 1 : gm);
                nwb_lc_prefetch(tab + (size_t)(j - 2 - NWB_LC_AHEAD) * pitch + (size_t)gm * 16);
            }
            /* --- the window's nibbles of this row: nibble index base - 1 + k --- */
            unsigned x[NX];
            {
                unsigned w[12] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w};
                const int c = base - 1;
                const int ow = (c >> 3) & 3; /* first word inside the three groups */
                if (ow & 2) {
#pragma unroll
                    for (int i = 0; i < 10; i++) w[i] = w[i + 2];
                }
                if (ow & 1) {
#pragma unroll
                    for (int i = 0; i < 7; i++) w[i] = w[i + 1];
                }
                const unsigned s = 4u * (unsigned)(c & 7);
#pragma unroll
                for (int i = 0; i < NX; i++) x[i] = __funnelshift_r(w[i], w[i + 1], s);
                /* columns beyond the top string hold nothing (the fill kernels may have left codes there) */
                const int nv = A - base + 1; /* window cells with column <= A */
                if (nv < W) {
#pragma unroll
                    for (int i = 0; i < NX; i++) {
                        const int left = nv - 8 * i;
                        x[i] &= (left >= 8) ? 0xFFFFFFFFu : ((left <= 0) ? 0u : ((1u << (4 * left)) - 1u));
                    }
                }
            }
            /* --- the row: P right to left, flow into row j - 1 --- */
            unsigned long long d = 0ull, dgp = 0ull;
#pragma unroll
            for (int k = W - 1; k >= 0; k--) {
                const unsigned f = x[k >> 3] >> (4 * (k & 7));
                const unsigned long long P = inc[k] + d;
                d = nwb_lc_sel(f & 2u, P);
                const unsigned long long up = nwb_lc_sel(f & 4u, P);
                if (k == W - 1) {
                    if (up != 0ull) bailed = true; /* UP out of the window's right end */
                } else {
                    inc[k + 1] = up + dgp;
                }
                dgp = nwb_lc_sel(f & 1u, P);
            }
            inc[0] = dgp;
            if (d != 0ull) bailed = true; /* a LEFT run leaves the window */
            /* --- move the window --- */
            if (sh < 0) { /* window 8 columns to the left: the cells move right */
                unsigned long long lost = 0ull;
#pragma unroll
                for (int k = W - NWB_LC_SHIFT; k < W; k++) lost |= inc[k];
                if (lost != 0ull) bailed = true;
#pragma unroll
                for (int k = W - 1; k >= NWB_LC_SHIFT; k--) inc[k] = inc[k - NWB_LC_SHIFT];
#pragma unroll
                for (int k = 0; k < NWB_LC_SHIFT; k++) inc[k] = 0ull;
            } else if (sh > 0) {
                unsigned long long lost = 0ull;
#pragma unroll
                for (int k = 0; k < NWB_LC_SHIFT; k++) lost |= inc[k];
                if (lost != 0ull) bailed = true;
#pragma unroll
                for (int k = 0; k < W - NWB_LC_SHIFT; k++) inc[k] = inc[k + NWB_LC_SHIFT];
#pragma unroll
                for (int k = W - NWB_LC_SHIFT; k < W; k++) inc[k] = 0ull;
            }
            base = nbase;
            j--;
            c0 = n0; c1 = n1; c2 = n2;
        }
        if (valid) {
            if (!active) {
                cp.out_count[p] = 1ull;
            } else if (bailed) {
                const unsigned pos = atomicAdd(cp.fb_count, 1u);
                cp.fb_list[pos] = p;
            } else {
                unsigned long long total = 0ull; /* what is left has arrived in row 0 */
#pragma unroll
                for (int k = 0; k < W; k++) total += inc[k];
                cp.out_count[p] = total;
            }
        }
    }
}
