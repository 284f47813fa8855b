/*
 * nwb_count_sparse.cuh -- the optimal-alignment count behind `-s`, swept BACKWARDS
 * from (A,B) over the cells that actually lie on optimal paths.
 *
 * get_solution_count() (computation.c:249-260; the reference enumerates every
 * alignment, needleman-wunsch.c:209-331) is the number of arrow paths from (A,B)
 * to (0,0).  nwb_count.cuh computes it with a dense forward DP over all A*B cells.
 * But with P(i,j) = number of arrow paths from (A,B) to (i,j),
 *     P(A,B) = 1,
 *     P(i,j) = [DIAG(i+1,j+1)] P(i+1,j+1) + [LEFT(i+1,j)] P(i+1,j) + [UP(i,j+1)] P(i,j+1),
 * only the cells reachable from (A,B) are non-zero: a band a few cells wide around
 * the optimal alignments (<= 26 columns on the BASELINE configs), and the count is
 * the flow that reaches the border (row 0 / column 0 have one forced arrow each,
 * computation.c:97-124).  Arithmetic is mod 2^64 like the dense DP, so a cell whose
 * value is 0 mod 2^64 contributes nothing and is dropped too; when no live cell is
 * left the sweep stops (on random DNA/protein the 2-adic valuation of P grows by
 * about one bit per dozen rows, so the live set dies ~1000 rows above (A,B)).
 *
 * One warp walks the rows j = B .. 1.  Lane l owns one unit of CPL consecutive columns
 * (CPL = 2: a byte of the arrow row, CPL = 8: a 32-bit word): the unit b == l (mod 32)
 * inside a window of 32 units (64 or 256 columns) that follows the live cells to the left.  Per row: P along the row
 * (right to left through runs of LEFT arrows: a local pass per lane, then carries
 * between lanes until none is left), then the flow into row j-1 (UP, DIAG).  Arrow
 * words are fetched one chunk of 8 rows ahead; a lane whose block has fallen off the
 * right end of the live band is re-assigned 32 blocks to the left for the NEXT chunk.
 * If the live band does not fit the window (or outruns the prefetch), the sweep gives
 * up (NWB_SPC_BAILED) and the dense forward sweep (nwb_count.cuh) runs instead; it is
 * launched behind this kernel in any case and returns at once when this one finished.
 */
#pragma once
#include "nwb_device.cuh"

#define NWB_SPC_CHUNK 8
/* NwbDevSummary.count_state */
#define NWB_SPC_NONE 0
#define NWB_SPC_DONE 1   /* summary->count is final (sparse sweep)          */
#define NWB_SPC_BAILED 2 /* live band too wide: the dense sweep must run     */

struct NwbSparseCountParams {
    const uint8_t *arrows; /* nibble table, B rows x pitch bytes (include/nwb.h layout) */
    size_t pitch;
    int A, B;
    unsigned long long *out_count; /* where the count goes                                  */
    int *out_state;                /* NWB_SPC_*                                             */
    unsigned *out_rows;            /* rows visited before the live set died (or NULL)       */
    int mode;                      /* tests: 1 = skip the 64-column attempt, 2 = skip the 256-column attempt */
    int min_col;                   /* 0 / 1: the whole table is here.  c > 1 (c - 1 a multiple of 256): `arrows` holds columns
                                    * c .. A only -- the last rank of a strip group; flow that would leave to the left of
                                    * column c makes the sweep give up instead of being counted at the border */
};

/* the CPL arrow nibbles of unit `unit` (CPL consecutive columns) in row j; 0 outside the table */
template <int CPL>
__device__ __forceinline__ unsigned nwb_spc_load(const uint8_t *arrows, size_t pitch, int j, int unit, unsigned colmask, int umin)
{
    if (j < 1 || unit < umin) return 0u;
    const int col0 = unit * CPL; /* 0-based first column of the unit */
    const unsigned *q = reinterpret_cast<const unsigned *>(arrows + (size_t)(j - 1) * pitch) + (col0 >> 3);
#ifdef NWB_EMU
    const unsigned w = *q;
#else
    const unsigned w = __ldca(q); /* a row's words sit in one or two 128-byte lines */
#endif
    return (w >> (4 * (col0 & 7))) & colmask;
}

__device__ __forceinline__ unsigned long long nwb_spc_sel(unsigned flag, unsigned long long v) { return flag ? v : 0ull; }

/* One pair: the warp's 32 lanes, CPL columns per lane (window of 32 * CPL columns).  Returns NWB_SPC_DONE /
 * NWB_SPC_BAILED (uniform); *count and *rows valid in every lane. */
template <int CPL>
__device__ __forceinline__ int nwb_sparse_count_pair(const uint8_t *arrows, const size_t pitch, const int A, const int B,
                                                     const int min_col, const int lane, unsigned long long *count, unsigned *rows)
{
    const int umin = (min_col > 1) ? (min_col - 1) / CPL : 0; /* first unit that is present */
    const unsigned FULL = (CPL == 8) ? 0x77777777u : (0x77777777u & ((1u << (4 * CPL)) - 1u));
    const int rb0 = (A - 1) / CPL;                /* unit of column A */
    int myb = rb0 - ((rb0 - lane) & 31);          /* my unit: == lane (mod 32), inside [rb0-31, rb0]; < 0: none */
    const int right = (lane + 1) & 31;            /* the lane that owns unit myb + 1 (cyclically)               */
    auto colmask_of = [&](const int b) -> unsigned {
        if (b != rb0) return FULL;
        const int n = ((A - 1) % CPL) + 1;        /* cells of the last unit that are inside the table */
        return n >= CPL ? FULL : (FULL & ((1u << (4 * n)) - 1u));
    };

    unsigned long long inc[CPL], P[CPL];
    {
        /* P(A,B) = 1 (as selects: an indexed store would put inc[] into local memory) */
        const int kA = (myb == rb0) ? ((A - 1) % CPL) : -1;
#pragma unroll
        for (int k = 0; k < CPL; k++) inc[k] = (k == kA) ? 1ull : 0ull;
    }
    unsigned long long total = 0ull; /* flow that has reached the border */
    unsigned w[NWB_SPC_CHUNK], wn[NWB_SPC_CHUNK];
    int jtop = B;
#pragma unroll
    for (int t = 0; t < NWB_SPC_CHUNK; t++) w[t] = nwb_spc_load<CPL>(arrows, pitch, jtop - t, myb, colmask_of(myb), umin);
    int rlb = rb0; /* rightmost live unit */
    int state = NWB_SPC_NONE;
    unsigned nrows = 0;

    while (state == NWB_SPC_NONE) {
        /* next chunk: a lane right of the live band moves 32 units to the left */
        const int nextb = (myb > rlb) ? myb - 32 : myb;
        {
            const unsigned cm = colmask_of(nextb);
#pragma unroll
            for (int t = 0; t < NWB_SPC_CHUNK; t++) wn[t] = nwb_spc_load<CPL>(arrows, pitch, jtop - NWB_SPC_CHUNK - t, nextb, cm, umin);
        }
        const int nbb = __shfl_sync(NWB_FULL_MASK, myb, right);
        const bool adjacent = (nbb == myb + 1);
        bool bad = false;
#pragma unroll
        for (int t = 0; t < NWB_SPC_CHUNK; t++) {
            const int j = jtop - t;
            if (j < 1 || state != NWB_SPC_NONE) continue; /* (no break: the loop must unroll so that w[] stays in registers) */
            const unsigned x = w[t];
            /* P along the row, right to left: P(i) = inc(i) + [LEFT(i+1)] P(i+1) */
            unsigned long long d = 0ull;
#pragma unroll
            for (int k = CPL - 1; k >= 0; k--) {
                P[k] = inc[k] + d;
                d = nwb_spc_sel((x >> (4 * k + 1)) & 1u, P[k]);
            }
            for (;;) {
                if (myb == umin) { /* LEFT out of column 1: the border column -- or out of this rank's columns */
                    if (umin == 0) total += d;
                    else bad = bad || (d != 0ull);
                    d = 0ull;
                }
                const unsigned long long cin = __shfl_sync(NWB_FULL_MASK, d, right);
                /* a carry that arrives from a lane which does not own the unit to my right has left the window: it is
                 * dropped (so it cannot circulate) and the sweep is given up at the end of the chunk */
                const bool take = (cin != 0ull) && adjacent;
                bad = bad || (cin != 0ull && !adjacent);
                if (!__any_sync(NWB_FULL_MASK, take)) break;
                d = take ? cin : 0ull;
#pragma unroll
                for (int k = CPL - 1; k >= 0; k--) {
                    P[k] += d;
                    d = nwb_spc_sel((x >> (4 * k + 1)) & 1u, d);
                }
            }
            /* flow into row j-1: UP keeps the column, DIAG moves one to the left */
            unsigned long long dg = nwb_spc_sel(x & 1u, P[0]);
            if (myb == umin) { /* DIAG out of column 1 (or out of this rank's columns) */
                if (umin == 0) total += dg;
                else bad = bad || (dg != 0ull);
                dg = 0ull;
            }
            const unsigned long long dgin = __shfl_sync(NWB_FULL_MASK, dg, right);
            bad = bad || (dgin != 0ull && !adjacent);
#pragma unroll
            for (int k = 0; k < CPL - 1; k++)
                inc[k] = nwb_spc_sel((x >> (4 * k + 2)) & 1u, P[k]) + nwb_spc_sel((x >> (4 * k + 4)) & 1u, P[k + 1]);
            inc[CPL - 1] = nwb_spc_sel((x >> (4 * (CPL - 1) + 2)) & 1u, P[CPL - 1]) + dgin;
            nrows++;
            /* the votes that end the sweep are taken once per chunk: a few rows without live cells cost nothing, and a
             * chunk that lost a carry is thrown away anyway */
            if (t == NWB_SPC_CHUNK - 1 || j == 1) {
                bool live = false;
#pragma unroll
                for (int k = 0; k < CPL; k++) live = live || (inc[k] != 0ull);
                if (__any_sync(NWB_FULL_MASK, bad)) { state = NWB_SPC_BAILED; continue; }
                if (!__any_sync(NWB_FULL_MASK, live)) { state = NWB_SPC_DONE; continue; }
                /* rightmost live unit, for the re-assignment at the chunk boundary */
                int v = live ? myb : -1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const int u = __shfl_xor_sync(NWB_FULL_MASK, v, o);
                    v = u > v ? u : v;
                }
                rlb = v;
            }
        }
        if (state != NWB_SPC_NONE) break;
        jtop -= NWB_SPC_CHUNK;
        if (jtop < 1) { state = NWB_SPC_DONE; break; }
        myb = nextb;
#pragma unroll
        for (int t = 0; t < NWB_SPC_CHUNK; t++) w[t] = wn[t];
    }
    /* whatever is left has arrived in row 0 (or is zero) */
#pragma unroll
    for (int k = 0; k < CPL; k++) total += inc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(NWB_FULL_MASK, total, o);
    *count = total;
    *rows = nrows;
    return state;
}

/* First with 2 columns per lane (a 64-column window: a row costs ~1/3 of the 8-column form, and the live band of the
 * BASELINE inputs is at most 26 columns wide), then, if that window was too narrow, with 8 (256 columns). */
__global__ void __launch_bounds__(32, 1) nwb_sparse_count_kernel(const NwbSparseCountParams p)
{
    const int lane = threadIdx.x & 31;
    unsigned long long count = 0ull;
    unsigned rows = 0u, rows2 = 0u;
    int state = NWB_SPC_BAILED;
    if (!(p.mode & 1)) state = nwb_sparse_count_pair<2>(p.arrows, p.pitch, p.A, p.B, p.min_col, lane, &count, &rows);
    if (state != NWB_SPC_DONE && !(p.mode & 2)) {
        state = nwb_sparse_count_pair<8>(p.arrows, p.pitch, p.A, p.B, p.min_col, lane, &count, &rows2);
        rows += rows2;
    }
    if (lane == 0) {
        if (state == NWB_SPC_DONE) *p.out_count = count;
        if (p.out_rows) *p.out_rows = rows;
        *p.out_state = state;
    }
}
