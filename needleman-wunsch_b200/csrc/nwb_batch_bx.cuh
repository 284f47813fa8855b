/*
 * nwb_batch_bx.cuh -- batch of independent pairs, TWO pairs per warp ("bx").
 *
 * Same job as nwb_batch.cuh (the reference looped per pair: alloc_computation +
 * init_computation + compute_table_scores + free_computation, computation.c:51-214,
 * needleman-wunsch.c:583) and the same results, for batches whose top strings are at
 * most 256 characters and whose per-cell differences fit a nibble (2d + m <= 7,
 * nwb_hx_supported).
 *
 * nwb_batch.cuh packs two COLUMN BLOCKS of one pair into the 16-bit halves of a
 * register (64 virtual lanes on an anti-diagonal: 63 steps of skew for a table of
 * 256 rows) and turns differences into arrow codes with compare instructions.  A
 * batch has a better source of pairs of cells: two DIFFERENT PAIRS.  Here the low
 * halves of every register belong to pair 2q, the high halves to pair 2q + 1; lane l
 * owns columns 8l+1 .. 8l+8 of both pairs and works on row t - l + 1 at step t (31
 * steps of skew), one __shfl_up_sync per step carries the neighbour's v of both pairs.
 * The differences of a row are packed as nibbles with integer multiply-adds and the
 * three zero tests (DIAG <=> z == a, LEFT <=> u == 0, UP <=> v == 0: every tie keeps
 * its arrow, needleman-wunsch.c:485-503) are done on 8 cells per instruction exactly as
 * in nwb_fill_hx.cuh.  No boundary streams, no inter-warp synchronisation.
 *
 * Arrow rows are de-skewed through a per-warp shared-memory ring (slot = row mod 64,
 * one ring per pair) and leave with LDS.128 / STG.128, four whole 128-byte rows per
 * instruction.  Steps whose lanes are all strictly inside both tables run an
 * unchecked body; the first 31 steps (lanes still above row 1) and the steps around
 * and below the last rows run a checked one (row masks for the branch counter,
 * walk-table.c:108-120, and the capture of sum_i u(i,B) = r(A,B)).
 */
#pragma once
#include "nwb_fill_hx.cuh"
#include "nwb_batch.cuh"
#ifndef NWB_EMU
#include <cuda_fp16.h>
#endif

#define NWB_BX_WARPS 12
#define NWB_BX_MAX_A 256
#define NWB_BX_RING_ROWS 64
#define NWB_BX_RING_WORDS (NWB_BX_RING_ROWS * 32) /* per pair: 64 rows x 128 bytes */
#define NWB_BX_SPADF 32  /* side words in front of row 1: lane 31 starts 30 rows above the table */
#define NWB_BX_STAIL 72  /* ... and behind the longer side string: 31 steps of skew + 31 of block rounding + prefetch */
#define NWB_BX_SIDE_WORDS(maxB) ((size_t)(maxB) + NWB_BX_SPADF + NWB_BX_STAIL)
#define NWB_BX_SMEM_PER_WARP(maxB) (2 * NWB_BX_RING_WORDS * 4 + ((NWB_BX_SIDE_WORDS(maxB) * 4 + 15) / 16) * 16)

/* whether a batch runs this kernel (host and emulator harness share the rule) */
static inline bool nwb_bx_usable(const NwbPkConsts &pc, long long max_A, int max_B)
{
    return nwb_hx_supported(pc) && max_A <= NWB_BX_MAX_A && NWB_BX_SMEM_PER_WARP(max_B) <= 220 * 1024;
}

/* Characters travel as fp16 bit patterns 0x3C00 + c (1.0 + c/1024: distinct normal numbers), rows outside a
 * table as 0xFFFF (a NaN: equal to nothing), so that one HSET2.EQ on the half-precision pipe -- not the ALU pipe
 * the DPX instructions and the bit logic compete for -- yields 0xFFFF / 0 per pair, and one LOP3 selects
 * a_match / a_mis from it. */
#define NWB_BX_CHAR(c) (0x3C00u + (unsigned)(c))
#define NWB_BX_NOCHAR 0xFFFFu
__device__ __forceinline__ unsigned nwb_bx_eq_mask(const unsigned x, const unsigned y)
{
#ifdef NWB_EMU
    unsigned m = 0u;
    for (int h = 0; h < 2; h++) {
        const unsigned a = (x >> (16 * h)) & 0xFFFFu, b = (y >> (16 * h)) & 0xFFFFu;
        const bool nan_a = (a & 0x7C00u) == 0x7C00u && (a & 0x03FFu), nan_b = (b & 0x7C00u) == 0x7C00u && (b & 0x03FFu);
        const bool zero = !(a & 0x7FFFu) && !(b & 0x7FFFu);
        if (!nan_a && !nan_b && (a == b || zero)) m |= 0xFFFFu << (16 * h);
    }
    return m;
#else
    return __heq2_mask(*reinterpret_cast<const __half2 *>(&x), *reinterpret_cast<const __half2 *>(&y));
#endif
}

struct NwbBxState {
    unsigned tpw[8]; /* top characters of my 8 columns as fp16 patterns (pair 2q | pair 2q+1)   */
    unsigned u[8];   /* u of my columns in the row above                                        */
    unsigned send;   /* v of my last column in the row I just finished (rows above row 1: BIG) */
    unsigned nu_a;   /* u of columns 0..3 / 4..7 in the row above, one nibble per column       */
    unsigned nu_b;
};

/* One step of one lane: row j = t - lane + 1 of both pairs, 8 cells each.  sp = the side characters of
 * row j as fp16 patterns (pair 2q | pair 2q+1).  roff = byte offset of row j's slot in the rings.
 * MODE 0: every lane is strictly inside both tables (1 <= j < B);  MODE 1: the first steps of a pair of
 * pairs, `live` = this lane has reached row 1 (j <= 0 otherwise; j < B for every lane);  MODE 2: j may lie
 * anywhere -- row masks for the branch counter and the capture of the bottom row. */
template <int MODE, int RING_ROWS = NWB_BX_RING_ROWS>
__device__ __forceinline__ void nwb_bx_step(NwbBxState &st, const NwbPkConsts &pc, const int lane, const unsigned sp,
                                             const nwb_smem_addr ring_l, unsigned &roff, unsigned cm_a, unsigned cm_b,
                                             const bool live, const int j, const int B0, const int B1,
                                             unsigned &br0, unsigned &br1, unsigned &rs0, unsigned &rs1)
{
    const unsigned nl0 = lane ? 1u : 0u;
    const unsigned AM = pc.TT1 - 0x00010001u; /* a_match in both halves */
    unsigned v = __shfl_up_sync(NWB_FULL_MASK, st.send, 1) * nl0; /* lane 0: column 0, r(0,j) = 0 */
    unsigned z[8], a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const unsigned eq = nwb_bx_eq_mask(st.tpw[k], sp);  /* 0xFFFF per half on a match */
        a[k] = (eq & AM) | (~eq & pc.AMIS);                 /* a_match or a_mis           */
        z[k] = __vimax3_s16x2(a[k], v, st.u[k]);
        const unsigned un = z[k] - v;
        const unsigned vn = z[k] - st.u[k];
        st.u[k] = un;
        v = vn;
    }
    st.send = v;
    /* the row's z, a and u as one nibble per column (word a: columns 0..3, word b: columns 4..7) */
    const unsigned Z4a = ((z[3] * 16u + z[2]) * 16u + z[1]) * 16u + z[0];
    const unsigned Z4b = ((z[7] * 16u + z[6]) * 16u + z[5]) * 16u + z[4];
    const unsigned A4a = ((a[3] * 16u + a[2]) * 16u + a[1]) * 16u + a[0];
    const unsigned A4b = ((a[7] * 16u + a[6]) * 16u + a[5]) * 16u + a[4];
    const unsigned NUa = ((st.u[3] * 16u + st.u[2]) * 16u + st.u[1]) * 16u + st.u[0];
    const unsigned NUb = ((st.u[7] * 16u + st.u[6]) * 16u + st.u[5]) * 16u + st.u[4];
    const unsigned p1a = ((A4a - Z4a + NWB_HX_B8) & NWB_HX_B8) | NUa; /* bit 3: z == a (DIAG); low bits: u */
    const unsigned p1b = ((A4b - Z4b + NWB_HX_B8) & NWB_HX_B8) | NUb;
    const unsigned zva = st.nu_a - Z4a + NWB_HX_B8;                   /* 8 - v: bit 3 set iff v == 0 (UP)  */
    const unsigned zvb = st.nu_b - Z4b + NWB_HX_B8;
    st.nu_a = NUa;
    st.nu_b = NUb;
    unsigned ta, tb;
    const unsigned ca = nwb_hx_code(p1a, zva, ta);
    const unsigned cb = nwb_hx_code(p1b, zvb, tb);
    nwb_sts<uint32_t>(ring_l + roff, 0, __byte_perm(ca, cb, 0x5410));                     /* pair 2q: my 8 cells of row j */
    nwb_sts<uint32_t>(ring_l + roff, RING_ROWS * 128, __byte_perm(ca, cb, 0x7632));       /* pair 2q+1                    */
    if ((RING_ROWS & (RING_ROWS - 1)) == 0) {
        roff = (roff + 128u) & (unsigned)(RING_ROWS * 128 - 1);
    } else {
        roff += 128u;
        if (roff == (unsigned)(RING_ROWS * 128)) roff = 0u;
    }
    if (MODE == 1) {
        cm_a = live ? cm_a : 0u;
        cm_b = live ? cm_b : 0u;
    }
    if (MODE == 2) {
        unsigned rm = 0u;
        if ((unsigned)(j - 1) < (unsigned)B0) rm |= 0x0000FFFFu;
        if ((unsigned)(j - 1) < (unsigned)B1) rm |= 0xFFFF0000u;
        /* bottom row of a pair: r(A,B) = sum_i u(i,B) over the pair's columns, from the nibble words */
        if (j >= 1 && (j == B0 || j == B1)) {
            const unsigned xa = NUa & ((cm_a >> 3) * 15u), xb = NUb & ((cm_b >> 3) * 15u);
            unsigned sb = (xa & 0x0F0F0F0Fu) + ((xa >> 4) & 0x0F0F0F0Fu) + (xb & 0x0F0F0F0Fu) + ((xb >> 4) & 0x0F0F0F0Fu);
            sb = (sb & 0x00FF00FFu) + ((sb >> 8) & 0x00FF00FFu);
            if (j == B0) rs0 += sb & 0xFFFFu;
            if (j == B1) rs1 += sb >> 16;
        }
        cm_a &= rm;
        cm_b &= rm;
    }
    ta &= cm_a;
    tb &= cm_b;
    br0 += (unsigned)__popc(__byte_perm(ta, tb, 0x5410));
    br1 += (unsigned)__popc(__byte_perm(ta, tb, 0x7632));
}

/* Rows (from, upto] of both pairs leave the rings: lane x moves 16 bytes (chunk x & 7) of row rb + (x >> 3). */
__device__ __forceinline__ void nwb_bx_flush(const unsigned *ring, const int from, const int upto, const int B0, const int B1,
                                              uint8_t *tab0, uint8_t *tab1, const int lane)
{
    const int sub = lane >> 3, chunk = lane & 7;
    for (int rb = from + 1; rb <= upto; rb += 4) {
        const int r = rb + sub;
        if (r <= upto) {
            const unsigned *src = ring + (r & (NWB_BX_RING_ROWS - 1)) * 32 + chunk * 4;
            const size_t dst = (size_t)(r - 1) * 128 + (size_t)chunk * 16;
            if (r <= B0) *reinterpret_cast<uint4 *>(tab0 + dst) = *reinterpret_cast<const uint4 *>(src);
            if (r <= B1) *reinterpret_cast<uint4 *>(tab1 + dst) = *reinterpret_cast<const uint4 *>(src + NWB_BX_RING_WORDS);
        }
    }
}

__global__ void __launch_bounds__(32 * NWB_BX_WARPS, 1) nwb_batch_bx_kernel(const NwbBatchParams bp, const NwbPkConsts pc)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long gwarp = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    unsigned *ring = reinterpret_cast<unsigned *>(NWB_SMEM_BASE() + (size_t)warp * NWB_BX_SMEM_PER_WARP(bp.max_B));
    unsigned *sidew = ring + 2 * NWB_BX_RING_WORDS;
    const nwb_smem_addr ring_l = nwb_smem_address(reinterpret_cast<unsigned char *>(ring + lane));
    const unsigned *side_l = sidew + NWB_BX_SPADF + 1 - lane; /* side_l[t] = the word of my row at step t */

    for (long long q = gwarp; 2 * q < bp.n_pairs; q += nwarps) {
        const long long p0 = 2 * q, p1 = 2 * q + 1;
        const bool have1 = p1 < bp.n_pairs;
        const long long t0o = bp.top_off[p0], s0o = bp.side_off[p0];
        const long long t1o = have1 ? bp.top_off[p1] : 0, s1o = have1 ? bp.side_off[p1] : 0;
        int A0 = (int)(bp.top_off[p0 + 1] - t0o), B0 = (int)(bp.side_off[p0 + 1] - s0o);
        int A1 = have1 ? (int)(bp.top_off[p1 + 1] - t1o) : 0, B1 = have1 ? (int)(bp.side_off[p1 + 1] - s1o) : 0;
        if (A0 == 0 || B0 == 0) { /* borders only (computation.c:97-124) */
            if (lane == 0) {
                bp.out_score[p0] = (A0 == 0) ? -B0 * bp.d : -A0 * bp.d;
                if (bp.out_branch) bp.out_branch[p0] = 0u;
            }
            A0 = 0; B0 = 0;
        }
        if (have1 && (A1 == 0 || B1 == 0)) {
            if (lane == 0) {
                bp.out_score[p1] = (A1 == 0) ? -B1 * bp.d : -A1 * bp.d;
                if (bp.out_branch) bp.out_branch[p1] = 0u;
            }
            A1 = 0; B1 = 0;
        }
        const int Bmax = B0 > B1 ? B0 : B1;
        if (Bmax == 0) continue;
        /* rows every non-empty pair of the warp has */
        const int Bmin = (B0 == 0) ? B1 : ((B1 == 0) ? B0 : (B0 < B1 ? B0 : B1));

        __syncwarp(); /* the previous pairs' reads of sidew and the rings are done */
        const int nwords = Bmax + NWB_BX_SPADF + NWB_BX_STAIL;
        for (int e = lane; e < nwords; e += 32) {
            const int j = e - NWB_BX_SPADF;
            unsigned lo = NWB_BX_NOCHAR, hi = NWB_BX_NOCHAR;
            if (j >= 1 && j <= B0) lo = NWB_BX_CHAR(bp.sides[s0o + j - 1]);
            if (j >= 1 && j <= B1) hi = NWB_BX_CHAR(bp.sides[s1o + j - 1]);
            sidew[e] = lo | (hi << 16);
        }
        NwbBxState st;
        const int c0 = 8 * lane; /* my first column, 0-based */
        int na0 = A0 - c0, na1 = A1 - c0;
        na0 = na0 < 0 ? 0 : (na0 > 8 ? 8 : na0);
        na1 = na1 < 0 ? 0 : (na1 > 8 ? 8 : na1);
        unsigned cm_a = 0u, cm_b = 0u; /* bit 3 of the nibbles whose column is inside the pair's table */
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const unsigned lo = (k < na0) ? (unsigned)bp.tops[t0o + c0 + k] : 0u;
            const unsigned hi = (k < na1) ? (unsigned)bp.tops[t1o + c0 + k] : 0u;
            st.tpw[k] = NWB_BX_CHAR(lo) | (NWB_BX_CHAR(hi) << 16);
            st.u[k] = 0u;
            const unsigned bit = 8u << (4 * (k & 3));
            const unsigned m = ((k < na0) ? bit : 0u) | ((k < na1) ? (bit << 16) : 0u);
            if (k < 4) cm_a |= m;
            else cm_b |= m;
        }
        st.send = NWB_PK_BIG * 0x00010001u;
        st.nu_a = 0u;
        st.nu_b = 0u;
        uint8_t *tab0 = bp.arrows + bp.arrow_off[p0];
        uint8_t *tab1 = have1 ? bp.arrows + bp.arrow_off[p1] : tab0;
        unsigned br0 = 0u, br1 = 0u, rs0 = 0u, rs1 = 0u;
        unsigned roff = (unsigned)((1 - lane) & (NWB_BX_RING_ROWS - 1)) * 128u; /* slot of row 1 - lane */
        __syncwarp();

        int flushed = 0;
        const int nsteps = Bmax + 31; /* lane 31 finishes row Bmax at step Bmax + 30 */
        int t = 0;
        /* head: steps 0 .. 30, lanes l > t are still above row 1 (lane 0 reaches row 31) */
        if (Bmin > 31) {
#pragma unroll 1
            for (; t < 31; t++)
                nwb_bx_step<1>(st, pc, lane, side_l[t], ring_l, roff, cm_a, cm_b, t >= lane, 0, B0, B1, br0, br1, rs0, rs1);
        } else {
#pragma unroll 1
            for (; t < 31; t++)
                nwb_bx_step<2>(st, pc, lane, side_l[t], ring_l, roff, cm_a, cm_b, true, t - lane + 1, B0, B1, br0, br1, rs0, rs1);
        }
        while (t < nsteps) {
            /* steps t .. t+31: lane 0 reaches row t + 32, lane 31 starts at row t - 30 >= 1 */
            if (t + 32 < Bmin) {
#pragma unroll 1
                for (int sub = 0; sub < 4; sub++) {
                    unsigned sw[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) sw[i] = side_l[t + 8 * sub + i];
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        nwb_bx_step<0>(st, pc, lane, sw[i], ring_l, roff, cm_a, cm_b, true, 0, B0, B1, br0, br1, rs0, rs1);
                }
            } else {
#pragma unroll 1
                for (int i = 0; i < 32; i++)
                    nwb_bx_step<2>(st, pc, lane, side_l[t + i], ring_l, roff, cm_a, cm_b, true, t + i - lane + 1, B0, B1,
                                   br0, br1, rs0, rs1);
            }
            t += 32;
            /* rows <= t - 31 are complete in every lane */
            int upto = t - 31;
            if (upto > Bmax) upto = Bmax;
            __syncwarp();
            nwb_bx_flush(ring, flushed, upto, B0, B1, tab0, tab1, lane);
            flushed = upto;
            __syncwarp();
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rs0 += __shfl_xor_sync(NWB_FULL_MASK, rs0, o);
            rs1 += __shfl_xor_sync(NWB_FULL_MASK, rs1, o);
            br0 += __shfl_xor_sync(NWB_FULL_MASK, br0, o);
            br1 += __shfl_xor_sync(NWB_FULL_MASK, br1, o);
        }
        if (lane == 0) {
            /* score(A,B) = sum_i u(i,B) - d*(A+B) */
            if (B0 > 0) {
                bp.out_score[p0] = (int)(rs0 - (unsigned)bp.d * (unsigned)(A0 + B0));
                if (bp.out_branch) bp.out_branch[p0] = br0;
            }
            if (B1 > 0) {
                bp.out_score[p1] = (int)(rs1 - (unsigned)bp.d * (unsigned)(A1 + B1));
                if (bp.out_branch) bp.out_branch[p1] = br1;
            }
        }
    }
}

/* ---------------------------------------------------------------------------------------------------------
 * Uniform batches ("cx"): every pair is A x B with B a multiple of 32 (BASELINE config 4: 256 x 256).
 *
 * nwb_batch_bx_kernel drains its 32-lane anti-diagonal at the end of every pair of pairs and refills it for
 * the next one: 31 + 32 of the 287 steps of a 256-row table run a checked body with half of the lanes idle.
 * Here a warp sweeps its pairs of pairs BACK TO BACK as one tall table of n_q * B rows: lane l starts row 1 of
 * the next pair of pairs in the step after it finished row B of the current one.  Row 1 needs nothing from the
 * rows above but u = 0, so the hand-over is a per-lane reset at step n*B + l: capture sum_i u(i,B) and the
 * branch counter of the finished pair of pairs, clear u, take the next top characters (prefetched into
 * registers a whole table earlier).  Every step of every block runs the unchecked body; the 32 steps in which
 * the lanes reset one after the other ("transition block") add a short divergent branch.  Side characters are
 * double-buffered in shared memory (the next pair of pairs' words are written right after a transition block),
 * arrow rows leave the rings in aligned groups of 32 rows, which never straddle two tables.
 * --------------------------------------------------------------------------------------------------------- */
/* WARPS = 12: rings of 64 rows, flushed every 32 steps; WARPS = 16: rings of 48 rows (32 rows of lane skew + 16),
 * flushed every 16 steps, 128 registers per thread. */
#define NWB_CX_RING_ROWS(WARPS) ((WARPS) == 16 ? 48 : 64)
#define NWB_CX_SMEM_PER_WARP(B, WARPS) (2 * NWB_CX_RING_ROWS(WARPS) * 128 + (((size_t)(B) * 2 * 4 + 15) / 16) * 16)

static inline bool nwb_cx_usable(const NwbPkConsts &pc, bool uniform, long long A, int B, int warps = NWB_BX_WARPS)
{
    return uniform && nwb_hx_supported(pc) && A >= 1 && A <= NWB_BX_MAX_A && B >= 64 && B % 32 == 0 &&
           NWB_CX_SMEM_PER_WARP(B, warps) * (size_t)warps <= 226 * 1024;
}

__device__ __forceinline__ unsigned nwb_warp_sum(unsigned x)
{
#ifdef NWB_EMU
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(NWB_FULL_MASK, x, o);
    return x;
#else
    return __reduce_add_sync(NWB_FULL_MASK, x);
#endif
}

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1) nwb_batch_cx_kernel(const NwbBatchParams bp, const NwbPkConsts pc,
                                                                      const int A, const int B)
{
    constexpr int RR = NWB_CX_RING_ROWS(WARPS); /* ring rows */
    constexpr int BLK = RR - 32;                /* steps between flushes */
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long gwarp = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    const long long NQ = (bp.n_pairs + 1) / 2; /* pairs of pairs in the batch */
    if (gwarp >= NQ) return;
    const int n_q = (int)((NQ - gwarp + nwarps - 1) / nwarps); /* ... in this warp's chain */
    unsigned *ring = reinterpret_cast<unsigned *>(NWB_SMEM_BASE() + (size_t)warp * NWB_CX_SMEM_PER_WARP(B, WARPS));
    unsigned *sidew = ring + 2 * RR * 32; /* [2][B]: chain element n reads buffer n & 1 */
    const nwb_smem_addr ring_l = nwb_smem_address(reinterpret_cast<unsigned char *>(ring + lane));
    const size_t tab_bytes = (size_t)128 * (size_t)B;

    int na = A - 8 * lane;
    na = na < 0 ? 0 : (na > 8 ? 8 : na);
    unsigned cm_a = 0u, cm_b = 0u; /* bit 3 of the nibbles whose column is inside the tables (both halves) */
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const unsigned bit = (k < na) ? ((8u << (4 * (k & 3))) * 0x00010001u) : 0u;
        if (k < 4) cm_a |= bit;
        else cm_b |= bit;
    }
    const unsigned nm_a = (cm_a >> 3) * 15u, nm_b = (cm_b >> 3) * 15u; /* the same as whole nibbles */

    /* pairs of chain element n; a missing partner (odd batch) is swept as a copy of the first pair and dropped */
    auto pair_of = [&](const int n, long long &p0, long long &p1) -> bool {
        const long long q = gwarp + (long long)n * nwarps;
        p0 = 2 * q;
        p1 = 2 * q + 1;
        const bool have1 = p1 < bp.n_pairs;
        if (!have1) p1 = p0;
        return have1;
    };
    auto load_tops = [&](const int n, unsigned (&tn)[8]) {
        long long p0, p1;
        pair_of(n, p0, p1);
        const uint8_t *t0 = bp.tops + bp.top_off[p0] + 8 * lane, *t1 = bp.tops + bp.top_off[p1] + 8 * lane;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const unsigned lo = (k < na) ? (unsigned)t0[k] : 0u, hi = (k < na) ? (unsigned)t1[k] : 0u;
            tn[k] = NWB_BX_CHAR(lo) | (NWB_BX_CHAR(hi) << 16);
        }
    };
    auto write_sides = [&](const int n) {
        long long p0, p1;
        pair_of(n, p0, p1);
        const uint8_t *s0 = bp.sides + bp.side_off[p0], *s1 = bp.sides + bp.side_off[p1];
        unsigned *dst = sidew + (n & 1) * B;
#pragma unroll 4
        for (int r = lane; r < B; r += 32) dst[r] = NWB_BX_CHAR(s0[r]) | (NWB_BX_CHAR(s1[r]) << 16);
    };

    NwbBxState st;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        st.tpw[k] = NWB_BX_CHAR(0) * 0x00010001u;
        st.u[k] = 0u;
    }
    st.send = 0u;
    st.nu_a = 0u;
    st.nu_b = 0u;
    unsigned tn[8];
    load_tops(0, tn);
    write_sides(0);
    unsigned br0 = 0u, br1 = 0u, rs0 = 0u, rs1 = 0u; /* rs*: unused by the unchecked step */
    unsigned roff = (unsigned)((1 - lane + RR) % RR) * 128u; /* slot of virtual row 1 - lane */
    __syncwarp();

    const int nb = B / 32;
    int flushed = 0; /* virtual rows 1 .. flushed are in memory */
    int t = 0;       /* steps done */
    auto flush_to = [&](const int upto) { /* whole groups of BLK virtual rows; a group lies inside one table */
        __syncwarp();
        while (flushed < upto) {
            const int n = flushed / B;
            long long p0, p1;
            const bool have1 = pair_of(n, p0, p1);
            const int r0 = flushed - n * B; /* local rows r0+1 .. r0+BLK */
            uint8_t *tab0 = bp.arrows + (size_t)p0 * tab_bytes, *tab1 = bp.arrows + (size_t)p1 * tab_bytes;
            const int sub = lane >> 3, chunk = lane & 7;
#pragma unroll
            for (int g = 0; g < BLK / 4; g++) {
                const int lr = r0 + 4 * g + sub; /* local row - 1 */
                const unsigned *src = ring + ((flushed + 4 * g + sub + 1) % RR) * 32 + chunk * 4;
                const size_t dst = (size_t)lr * 128 + (size_t)chunk * 16;
                const uint4 w0 = *reinterpret_cast<const uint4 *>(src);
                const uint4 w1 = *reinterpret_cast<const uint4 *>(src + RR * 32);
                *reinterpret_cast<uint4 *>(tab0 + dst) = w0;
                if (have1) *reinterpret_cast<uint4 *>(tab1 + dst) = w1;
            }
            flushed += BLK;
        }
        __syncwarp();
    };

    for (int n = 0; n <= n_q; n++) {
        /* ---- transition block: steps n*B .. n*B+31, lane l moves on to chain element n at step n*B + l ---- */
        const unsigned *bufn = sidew + (n & 1) * B, *bufp = sidew + ((n + 1) & 1) * B;
        unsigned rc0 = 0u, rc1 = 0u, bc0 = 0u, bc1 = 0u;
#pragma unroll 1
        for (int i = 0; i < 32; i++) {
            const int li = i - lane; /* my row of element n, minus 1; negative: still in element n - 1 */
            const unsigned sp = (li >= 0) ? bufn[li] : bufp[B + li];
            if (li == 0) {
                /* my last row of element n - 1 is behind me: r(A,B) = sum_i u(i,B), from the nibble words */
                const unsigned xa = st.nu_a & nm_a, xb = st.nu_b & nm_b;
                unsigned sb = (xa & 0x0F0F0F0Fu) + ((xa >> 4) & 0x0F0F0F0Fu) + (xb & 0x0F0F0F0Fu) + ((xb >> 4) & 0x0F0F0F0Fu);
                sb = (sb & 0x00FF00FFu) + ((sb >> 8) & 0x00FF00FFu);
                rc0 = sb & 0xFFFFu;
                rc1 = sb >> 16;
                bc0 = br0;
                bc1 = br1;
                br0 = 0u;
                br1 = 0u;
                st.nu_a = 0u;
                st.nu_b = 0u;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    st.u[k] = 0u; /* row 0 of the new tables: r(i,0) = 0 */
                    st.tpw[k] = tn[k];
                }
            }
            nwb_bx_step<0, RR>(st, pc, lane, sp, ring_l, roff, cm_a, cm_b, true, 0, B, B, br0, br1, rs0, rs1);
            if (BLK < 32 && i == BLK - 1) flush_to(t + BLK - 32);
        }
        t += 32;
        flush_to(t - 32);
        if (n >= 1) { /* element n - 1 is complete in every lane */
            rc0 = nwb_warp_sum(rc0);
            rc1 = nwb_warp_sum(rc1);
            bc0 = nwb_warp_sum(bc0);
            bc1 = nwb_warp_sum(bc1);
            long long p0, p1;
            const bool have1 = pair_of(n - 1, p0, p1);
            if (lane == 0) {
                /* score(A,B) = sum_i u(i,B) - d*(A+B) */
                const unsigned off = (unsigned)bp.d * (unsigned)(A + B);
                bp.out_score[p0] = (int)(rc0 - off);
                if (bp.out_branch) bp.out_branch[p0] = bc0;
                if (have1) {
                    bp.out_score[p1] = (int)(rc1 - off);
                    if (bp.out_branch) bp.out_branch[p1] = bc1;
                }
            }
        }
        if (n == n_q) break;
        if (n + 1 < n_q) {
            write_sides(n + 1); /* its buffer was last read in the transition block that just ended */
            load_tops(n + 1, tn);
        }
        __syncwarp();
        /* ---- the rest of element n: every lane strictly inside the tables ---- */
        const unsigned *side_l = bufn + 32 - lane; /* side_l[s] = my row's word at step n*B + 32 + s */
#pragma unroll 1
        for (int blk = 32 / BLK; blk < nb * (32 / BLK); blk++) {
#pragma unroll 1
            for (int sub = 0; sub < BLK / 8; sub++) {
                unsigned sw[8];
#pragma unroll
                for (int i = 0; i < 8; i++) sw[i] = side_l[i];
                side_l += 8;
#pragma unroll
                for (int i = 0; i < 8; i++)
                    nwb_bx_step<0, RR>(st, pc, lane, sw[i], ring_l, roff, cm_a, cm_b, true, 0, B, B, br0, br1, rs0, rs1);
            }
            t += BLK;
            flush_to(t - 32);
        }
    }
    flush_to(n_q * B);
}
