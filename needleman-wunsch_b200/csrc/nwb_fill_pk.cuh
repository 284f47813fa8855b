/*
 * nwb_fill_pk.cuh -- packed 16x2 "difference" fill: two cells per instruction.
 *
 * Same results as the general kernel (nwb_fill_i32.cuh) -- i.e. as the
 * reference's score_cell() (needleman-wunsch.c:418-510) -- for scoring schemes
 * whose per-cell differences are small (nwb_pk_supported()).  Selected
 * automatically; everything else runs the int32 kernel.
 *
 * Formulation.  With r(i,j) = score(i,j) + d*(i+j) the gap penalty disappears:
 *     r(i,j) = max( r(i-1,j-1) + a(i,j),  r(i,j-1),  r(i-1,j) ),
 *     a = 2d + m on a match, 2d - k on a mismatch,  r = 0 on both borders.
 * Only differences are carried:  u = r(i,j) - r(i-1,j),  v = r(i,j) - r(i,j-1),
 *     z = max3(a, vL, uU)        (vL = v of the left cell, uU = u of the upper cell)
 *     u = z - vL,   v = z - uU
 * and the reference's tie rule "every candidate equal to the maximum gets an
 * arrow" (needleman-wunsch.c:485-503) becomes three zero tests:
 *     DIAG <=> z - a == 0,   LEFT <=> u == 0,   UP <=> v == 0.
 * All values live in [0, 2d+m], so two cells are packed in the 16-bit halves of
 * one register and processed by one DPX instruction each:
 *     VIADDMNMX.S16x2 (a from the characters), VIMNMX3.S16x2 (z), VIMNMX.U16x2
 *     (zero tests).  The optimal score is recovered as sum_i u(i,B) - d*(A+B).
 *
 * Mapping.  A strip is 64*K columns (K = 1, 2 or 4).  Lane l owns two column
 * blocks of K columns: block 2l in the low halves, block 2l+1 in the high
 * halves; the high half runs one row behind the low half and lane l+1 one row
 * behind lane l's high half (64 virtual lanes on an anti-diagonal).  Per row
 * step one __shfl_up_sync carries {v of my last column, side character} to the
 * right neighbour.  Strips are chained exactly as in the general kernel:
 * persistent warps, cyclic strip assignment, boundary stream + progress word
 * (st.release / ld.acquire), optional peer-GPU inbox/outbox.
 *
 * Rows above the table (j <= 0) need no predication: a virtual row fed with
 * vL = BIG keeps u = 0 and hands BIG on to the right, so lanes that have not
 * reached row 1 yet just idle on harmless values.
 */
#pragma once
#include "nwb_device.cuh"

#define NWB_PK_WARPS 4      /* default warps per block: one per SM sub-partition */
#define NWB_PK_MAX_WARPS 8
#define NWB_PK_RING_ROWS 128
#define NWB_PK_BIG 0x7FFFu

struct NwbPkConsts {
    int a_match, a_mis, c; /* a_match = 2d+m, a_mis = 2d-k, c = m+k            */
    int shift;             /* characters are pre-shifted by this many bits     */
    unsigned TT1;          /* (a_match + 1) in both halves                     */
    unsigned AMIS;         /* a_mis in both halves                             */
};

/* The packed path needs 0 <= a_mis <= a_match (mismatch never beats match, the
 * mismatch diagonal is not negative) and a match/mismatch gap c that fits the
 * character trick: (top ^ side) << shift is 0 on a match and >= 2^shift >= c
 * otherwise, and a_match - 255 * 2^shift must not wrap a signed 16-bit half. */
static inline bool nwb_pk_supported(int m, int k, int d, NwbPkConsts *pc)
{
    const long long am = 2LL * d + m, ax = 2LL * d - k, c = (long long)m + k;
    if (ax < 0 || am < ax || am > 4000 || c > 128) return false;
    int sh = 0;
    while ((1LL << sh) < c) sh++;
    if (sh > 7) return false;
    if (pc) {
        pc->a_match = (int)am;
        pc->a_mis = (int)ax;
        pc->c = (int)c;
        pc->shift = sh;
        pc->TT1 = (unsigned)(am + 1) * 0x00010001u;
        pc->AMIS = (unsigned)ax * 0x00010001u;
    }
    return true;
}

/* Columns per half-lane.  Every strip boundary costs a pipeline hop of ~72 row
 * steps (64 virtual lanes of skew plus the hand-off) while every column per lane
 * only lengthens a step; measured on B200 (10k, 30k, 100k squares; 1 and 2 GPUs)
 * the widest strip, K = 4 (256 columns), wins everywhere: on 2 GPUs K = 4 / 2 / 1
 * gave 972 / 940 / 677 GCUPS at 100k x 100k. */
static inline int nwb_pk_choose_k(int A, int B, int world)
{
    (void)A;
    (void)B;
    (void)world;
    return 4;
}

#define NWB_PK_SMEM_BYTES(K, R, WARPS) ((size_t)(WARPS) * (NWB_PK_RING_ROWS * (R) * (32 * (K)) + 256))

template <int K>
struct NwbPkStage;
template <>
struct NwbPkStage<4> { typedef uint32_t T; };
template <>
struct NwbPkStage<2> { typedef uint16_t T; };
template <>
struct NwbPkStage<1> { typedef uint8_t T; };

/* Boundary stream words validate themselves.  One 32-bit word per row GROUP (R
 * rows): R = 1: {bit 31: valid, bits 16..30: v};  R = 2: {bit 31: valid,
 * bits 16..30: v of the group's second row, bit 15: valid, bits 0..14: v of its
 * first row}, v = the strip's last-column vertical difference in that row.  The
 * buffers are zeroed before every fill; the consumer reads the words with relaxed
 * (L2) loads a few steps ahead of use and re-polls only if a word is not valid
 * yet, so the hand-off needs neither a separate flag nor fences: every word is
 * written and read atomically and is all the consumer needs for that row group. */
#define NWB_PK_VALID 0x80000000u
#define NWB_PK_SUB 8 /* stream words are fetched 8 row groups at a time, one sub-block ahead */
/* words in front of group 0 of every boundary stream: lane 31 publishes from step 0 on, 63 row
 * groups before its own group 0, without a range test (the words land in this padding) */
#define NWB_PK_BPAD 64
/* The side string is pre-shifted and complemented once per fill into a uint16
 * array padded on both sides, so the row loop never range-checks its index:
 * side_pre[j + NWB_PK_SPAD] = ~(side[j-1] << shift) & 0xFFFF for 1 <= j <= B.
 * NWB_PK_SPAD is odd so that row 1 + 2*n sits at an even (4-byte aligned) index. */
#define NWB_PK_SPAD 257
#define NWB_PK_SPRE_LEN(B) ((size_t)(B) + NWB_PK_SPAD + 640)

__global__ void nwb_pk_prep_side_kernel(const uint8_t *side, int B, int shift, uint16_t *side_pre)
{
    const size_t n = NWB_PK_SPRE_LEN(B);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const long long j = (long long)i - NWB_PK_SPAD;
        unsigned v = 0xFFFFu;
        if (j >= 1 && j <= B) v = (~((unsigned)side[j - 1] << shift)) & 0xFFFFu;
        side_pre[i] = (uint16_t)v;
    }
}

/* staging ring: one slot per row GROUP (slot g & 127, R sub-rows of 32*K bytes);
 * lane l drops its 2*K nibbles of a row at byte l*K of the sub-row, so a finished
 * row is contiguous and the flush moves it with LDS.128 / STG.128. */
#define NWB_PK_SUBROW_BYTES(K) (32 * (K))
#define NWB_PK_SLOT_BYTES(K, R) ((R) * NWB_PK_SUBROW_BYTES(K))

template <int K, int R>
struct NwbPkState {
    unsigned tpw[K];      /* pre-shifted top characters of my columns (low block | high block) */
    unsigned u[K];        /* u of my columns in the row above                                   */
    unsigned vlast[R];    /* v of my last columns per sub-row (virtual rows: BIG)               */
    unsigned sp[R];       /* ~(side char << shift) per sub-row (low half row | high half row)   */
    unsigned acc_prev[R]; /* low-half arrow nibbles of the previous step per sub-row            */
    unsigned send;        /* v of my HIGH block's last column for the R rows just done          */
};

__device__ __forceinline__ void nwb_st_relaxed_sys_pred(uint32_t *p, unsigned v, bool pred)
{
#ifdef NWB_EMU
    if (pred) *(volatile uint32_t *)p = v;
#else
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.relaxed.sys.global.u32 [%0], %1;\n\t}"
                 ::"l"(p), "r"(v), "r"((unsigned)pred));
#endif
}

/* ---- fused 64-bit path count (NWB_WANT_COUNT; replaces the enumeration behind
 * get_solution_count(), computation.c:249, by cnt = [diag]cnt_diag + [left]cnt_left + [up]cnt_up
 * in uint64 wrap-around arithmetic).  Counts cannot share a register, so each half keeps its
 * own uint64 per column; the three tie tests come as predicate pairs straight out of
 * VIMNMX.U16x2 (x <= 0 per half), the adds are predicated IADD3/IADD3.X pairs. */
template <int K, int R>
struct NwbPkCnt {
    unsigned long long cnt[K][2];   /* counts of my columns in the row above, [low block, high block] */
    unsigned long long clast[R][2]; /* counts of my blocks' last columns, per sub-row of the last step */
    unsigned long long cdiag[2];    /* count left of my block in the last row of the previous step     */
};

/* count stream words: 2 x 64 bit per row, each self-validating (bit 63):
 * word 0 = count bits 0..62, word 1 = count bit 63 */
#define NWB_PK_CVALID 0x8000000000000000ull

__device__ __forceinline__ void nwb_st_relaxed_sys_pred_u64(unsigned long long *p, unsigned long long v, bool pred)
{
#ifdef NWB_EMU
    if (pred) *(volatile unsigned long long *)p = v;
#else
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.relaxed.sys.global.u64 [%0], %1;\n\t}"
                 ::"l"(p), "l"(v), "r"((unsigned)pred));
#endif
}
__device__ __forceinline__ unsigned long long nwb_ld_relaxed_u64(const unsigned long long *p, bool sys)
{
#ifdef NWB_EMU
    (void)sys;
    return *(const volatile unsigned long long *)p;
#else
    unsigned long long v;
    if (sys) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
#endif
}

/* Per-strip constants of the row-range tests (one code body serves the head of a
 * strip, its bulk and its tail: the first and last 63 steps sit on the
 * strip-to-strip critical path and must not run through cold, separately
 * unrolled code). */
template <int R>
struct NwbPkRange {
    unsigned gcnt[R]; /* number of row groups whose sub-row r is inside the table (row <= B) */
    int capg;         /* the row group that contains row B */
    int rB;           /* ... and row B's sub-row in it      */
    unsigned ngroups;
};

/* One row step of one lane: R rows x 2*K cells.  At step s lane l's low block is
 * on row group s - 2l, its high block on group s - 2l - 1 (= g_hi); group g holds
 * rows R*g+1 .. R*g+R.  out_w + g_idx addresses the stream word of group g_hi
 * (g_idx is a compile-time constant against a rebased pointer in the unrolled loop).
 * LEAN: every lane is strictly inside rows 1..B-1 for the whole block, so the
 * row-range predicates and the bottom-row capture are compiled out (the bulk of a
 * strip).  The checked variant serves the first and last 63 steps; it is kept
 * small (4 steps unrolled) because it runs once per strip from a cold I-cache
 * and sits on the strip-to-strip critical path. */
template <int K, int R, bool LEAN, bool COUNT>
__device__ __forceinline__ void nwb_pk_step(NwbPkState<K, R> &st, const NwbPkConsts &pc, const NwbPkRange<R> &rg,
                                             const unsigned bq, const int t, const int lane, const int g_idx,
                                             const int g_hi, const int A, const int col_lo, const int col_hi,
                                             const unsigned chars, const nwb_smem_addr slot, uint32_t *out_w,
                                             const bool pub31, unsigned &rs32, NwbPkCnt<K, R> &cc,
                                             const unsigned long long *cstage, unsigned long long *out_c,
                                             unsigned long long &cfinal)
{
    typedef typename NwbPkStage<K>::T stage_t;
    const unsigned ONE = 0x00010001u;
    /* ---- left inputs: from my left neighbour lane; lane 0 from the (validated) stream word */
    unsigned recv = __shfl_up_sync(NWB_FULL_MASK, st.send, 1);
    const unsigned b = __shfl_sync(NWB_FULL_MASK, bq, t);
    if (lane == 0) recv = b;
    unsigned vL[R];
    if (R == 2) {
        vL[0] = __byte_perm(recv, st.vlast[0], 0x5410); /* lo <- neighbour's row 0, hi <- my low block's row 0 */
        vL[R - 1] = __byte_perm(recv, st.vlast[R - 1], 0x5432);
        st.sp[0] = __byte_perm(chars, st.sp[0], 0x5410);
        st.sp[R - 1] = __byte_perm(chars, st.sp[R - 1], 0x5432);
    } else {
        vL[0] = __byte_perm(recv, st.vlast[0], 0x5432);
        st.sp[0] = __byte_perm(chars, st.sp[0], 0x5410);
    }
    /* counts entering my blocks from the left, per sub-row: low block <- neighbour's high block
     * (lane 0: the staged stream counts), high block <- my own low block of the previous step */
    unsigned long long cL0[R][2];
    if (COUNT) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            unsigned long long rc = __shfl_up_sync(NWB_FULL_MASK, cc.clast[r][1], 1);
            if (lane == 0) rc = cstage[t * R + r];
            cL0[r][0] = rc;
            cL0[r][1] = cc.clast[r][0];
        }
    }
    /* ---- the cells: row by row, column by column (the scheduler finds the wavefront) */
    unsigned acc[R];
    unsigned uafter[R][K]; /* u right after sub-row r (only read by the rare bottom-row capture) */
#pragma unroll
    for (int r = 0; r < R; r++) {
        unsigned v = vL[r];
        unsigned code[K];
        unsigned long long cl[2] = {0ull, 0ull}, cd[2] = {0ull, 0ull};
        if (COUNT) {
            cl[0] = cL0[r][0]; cl[1] = cL0[r][1];
            cd[0] = (r == 0) ? cc.cdiag[0] : cL0[r > 0 ? r - 1 : 0][0];
            cd[1] = (r == 0) ? cc.cdiag[1] : cL0[r > 0 ? r - 1 : 0][1];
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            const unsigned nx = st.tpw[k] ^ st.sp[r];                      /* -x'-1 per half   */
            const unsigned a = __viaddmax_s16x2(nx, pc.TT1, pc.AMIS);     /* a_match or a_mis */
            const unsigned z = __vimax3_s16x2(a, v, st.u[k]);
            const unsigned un = z - v;
            const unsigned vn = z - st.u[k];
            const unsigned td = z - a;
            const unsigned fd = __vminu2(td, ONE), fl = __vminu2(un, ONE), fu = __vminu2(vn, ONE);
            code[k] = fd + fl * 2u + fu * 4u; /* inverted: a set bit = NO arrow */
            if (COUNT) {
                /* arrow present = 1 per half; the conditional 64-bit adds become 32x64-bit
                 * multiply-adds on the FMA pipe (IMAD.WIDE.U32 + IMAD), no predicates, no selects */
                const unsigned ad = fd ^ ONE, al = fl ^ ONE, au = fu ^ ONE;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const unsigned xd = h ? (ad >> 16) : (ad & 0xFFFFu);
                    const unsigned xl = h ? (al >> 16) : (al & 0xFFFFu);
                    const unsigned xu = h ? (au >> 16) : (au & 0xFFFFu);
                    const unsigned long long n = (unsigned long long)xd * cd[h] + (unsigned long long)xl * cl[h] +
                                                 (unsigned long long)xu * cc.cnt[k][h];
                    cd[h] = cc.cnt[k][h];
                    cc.cnt[k][h] = n;
                    cl[h] = n;
                }
            }
            st.u[k] = un;
            uafter[r][k] = un;
            v = vn;
        }
        st.vlast[r] = v;
        if (COUNT) {
            cc.clast[r][0] = cl[0];
            cc.clast[r][1] = cl[1];
            if (!LEAN && r == rg.rB) { /* cells[M-1][N-1]: the block that owns column A, when it is on row B */
                if (g_hi + 1 == rg.capg) {
#pragma unroll
                    for (int k = 0; k < K; k++)
                        if (col_lo + k == A) cfinal = cc.cnt[k][0];
                }
                if (g_hi == rg.capg) {
#pragma unroll
                    for (int k = 0; k < K; k++)
                        if (col_hi + k == A) cfinal = cc.cnt[k][1];
                }
            }
        }
        if (K == 4) acc[r] = (code[0] + code[1] * 16u) + (code[2] + code[3] * 16u) * 256u;
        else if (K == 2) acc[r] = code[0] + code[1] * 16u;
        else acc[r] = code[0];
    }
    if (COUNT) {
        cc.cdiag[0] = cL0[R - 1][0];
        cc.cdiag[1] = cL0[R - 1][1];
    }
    /* ---- outputs */
    st.send = (R == 2) ? __byte_perm(st.vlast[0], st.vlast[R - 1], 0x7632) : st.vlast[0];
#pragma unroll
    for (int r = 0; r < R; r++) {
        /* arrow codes of row R*g_hi+1+r: low block from the previous step, high block from this one;
         * skipped while the lane is above row 1 (g_hi < 0) or below row B */
        if (LEAN || (unsigned)g_hi < rg.gcnt[r]) {
            stage_t w;
            if (K == 4) w = (stage_t)__byte_perm(st.acc_prev[r], acc[r], 0x7610);
            else if (K == 2) w = (stage_t)((st.acc_prev[r] & 0xFFu) | ((acc[r] >> 8) & 0xFF00u));
            else w = (stage_t)((st.acc_prev[r] & 0xFu) | ((acc[r] >> 12) & 0xF0u));
            nwb_sts<stage_t>(slot, r * NWB_PK_SUBROW_BYTES(K), w);
        }
        st.acc_prev[r] = acc[r];
    }
    /* row B passes through a lane's low block in one step and through its high block in the
     * next: bottom row, r(A,B) = sum of u(i,B) */
    if (!LEAN && __builtin_expect((unsigned)(rg.capg - g_hi) <= 1u, 0)) {
        const unsigned half = (g_hi == rg.capg) ? 0xFFFF0000u : 0x0000FFFFu; /* which block is on row B now */
#pragma unroll
        for (int k = 0; k < K; k++) {
            unsigned x = uafter[0][k];
            if (R == 2 && rg.rB == 1) x = uafter[R - 1][k];
            unsigned m = 0u;
            if (col_lo + k <= A) m |= 0x0000FFFFu;
            if (col_hi + k <= A) m |= 0xFFFF0000u;
            rs32 += x & m & half; /* packed: each half sums at most K values below 2^13 */
        }
    }
    /* lane 31: the strip's last column for group g_hi, self-validating */
    const bool pub = pub31 && (LEAN || (unsigned)g_hi < rg.ngroups);
    nwb_st_relaxed_sys_pred(out_w + g_idx, st.send | ((R == 2) ? 0x80008000u : 0x80000000u), pub);
    if (COUNT) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            const unsigned long long cv = cc.clast[r][1];
            nwb_st_relaxed_sys_pred_u64(out_c + (g_idx * R + r) * 2, cv | NWB_PK_CVALID, pub);
            nwb_st_relaxed_sys_pred_u64(out_c + (g_idx * R + r) * 2 + 1, (cv >> 63) | NWB_PK_CVALID, pub);
        }
    }
}

/* branch counter (get_branch_count(), walk-table.c:133-147) of 32 flushed cells: a cell branches
 * when it has two or more arrows, i.e. at most one of its three inverted flags is set */
__device__ __forceinline__ unsigned nwb_pk_branches(const uint4 arrows, const uint4 colmask)
{
    const unsigned w[4] = {arrows.x, arrows.y, arrows.z, arrows.w};
    const unsigned m[4] = {colmask.x, colmask.y, colmask.z, colmask.w};
    unsigned n = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const unsigned x = w[q], x1 = x >> 1, x2 = x >> 2;
        const unsigned two = (x & x1) | (x & x2) | (x1 & x2); /* bit 0 of each nibble: >= 2 arrows */
        n += (unsigned)__popc(two & m[q]);
    }
    return n;
}

/* side characters of R consecutive rows: from the global side_pre array (read-only
 * path) or, in the batch kernel, from the warp's shared-memory copy */
template <int R, bool SMEMCH>
__device__ __forceinline__ unsigned nwb_pk_chars(const uint16_t *q)
{
    if (R == 2) return SMEMCH ? *reinterpret_cast<const unsigned *>(q) : nwb_ldg_u32(reinterpret_cast<const unsigned *>(q));
    return SMEMCH ? (unsigned)*q : (unsigned)nwb_ldg_u16(q);
}

template <int K, int R, bool SMEMCH, bool COUNT>
__device__ __forceinline__ bool nwb_pk_strip(const NwbStripParams &p, const NwbPkConsts &pc, const int c,
                                              unsigned char *stage_bytes, const int lane, long long &rsum,
                                              unsigned long long *cstage, unsigned long long &cfinal,
                                              const bool count_branches, unsigned &branches)
{
    typedef typename NwbPkStage<K>::T stage_t;
    const int A = p.A, B = p.B;
    const int W = 64 * K;
    const int col_lo = c * W + (2 * lane) * K + 1; /* first column (1-based) of the low block */
    const int col_hi = col_lo + K;                 /* ... of the high block                   */
    const unsigned ONE = 0x00010001u;
    const int SLOT = NWB_PK_SLOT_BYTES(K, R);
    const int ngroups = (B + R - 1) / R;

    NwbPkState<K, R> st;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const unsigned lo = (col_lo + k <= A) ? (unsigned)p.top[col_lo + k - 1] : 0u;
        const unsigned hi = (col_hi + k <= A) ? (unsigned)p.top[col_hi + k - 1] : 0u;
        st.tpw[k] = ((lo << pc.shift) | ((hi << pc.shift) << 16));
        st.u[k] = 0u;
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        st.vlast[r] = NWB_PK_BIG * ONE;
        st.sp[r] = 0xFFFFFFFFu;
        st.acc_prev[r] = 0u;
    }
    st.send = NWB_PK_BIG * ONE;
    /* counts: the border row/column and every virtual cell above the table count 1 path
     * (a virtual cell has only its LEFT arrow, so it copies the 1 handed in from the left) */
    NwbPkCnt<K, R> cc;
#pragma unroll
    for (int k = 0; k < K; k++) cc.cnt[k][0] = cc.cnt[k][1] = 1ull;
#pragma unroll
    for (int r = 0; r < R; r++) cc.clast[r][0] = cc.clast[r][1] = 1ull;
    cc.cdiag[0] = cc.cdiag[1] = 1ull;

    const int lc = c - p.strip_begin;
    const bool has_left = (c > 0);
    const bool left_remote = has_left && (lc == 0);
    const bool publish = (c + 1 < p.n_strips);
    const bool out_remote = publish && (c == p.strip_end - 1);
    uint32_t *out_w = (out_remote ? p.out_bnd_w : p.bnd_w + (size_t)lc * p.bpitch) + NWB_PK_BPAD;
    const uint32_t *in_w = nullptr;
    if (has_left) in_w = (left_remote ? p.in_bnd_w : p.bnd_w + (size_t)(lc - 1) * p.bpitch) + NWB_PK_BPAD;
    const bool pub31 = publish && (lane == 31) && !NWB_FAULT_INJECTED(p);
    const bool is_last = (c == p.n_strips - 1);
    /* count streams: 2 x uint64 per row, (BPAD + group) * R + sub-row */
    unsigned long long *out_c = nullptr;
    const unsigned long long *in_c = nullptr;
    if (COUNT) {
        out_c = (out_remote ? p.out_bnd_c : p.bnd_c + (size_t)lc * 2 * p.bpitch) + (size_t)NWB_PK_BPAD * R * 2;
        if (has_left) in_c = (left_remote ? p.in_bnd_c : p.bnd_c + (size_t)(lc - 1) * 2 * p.bpitch) + (size_t)NWB_PK_BPAD * R * 2;
        if (!has_left && lane < NWB_PK_SUB * R) cstage[lane] = 1ull; /* column 0 of the table */
        __syncwarp();
    }
    const bool nowait = NWB_DBG_BITS(p, 1) != 0;
    /* my low block's first row at step s is R*(s - 2*lane) + 1 */
    const uint16_t *sp_lane = p.side_pre + NWB_PK_SPAD + 1 - 2 * R * lane;

    /* stream words: lane i < 8 holds the word of row group 8*q + i of the current sub-block q;
     * the next sub-block's words are loaded one sub-block (8 steps) ahead into bq_next.
     * Column 0 of the table (no left strip): v(0,j) = 0. */
    const unsigned VMASK = (R == 2) ? 0x7FFF7FFFu : 0x7FFFFFFFu;
    unsigned bq = 0u, bq_next = 0u;
    if (has_left && lane < NWB_PK_SUB && lane < ngroups) bq_next = nwb_ld_relaxed_u32(in_w + lane, left_remote);

    const nwb_smem_addr lane_stage = nwb_smem_address(stage_bytes) + (unsigned)(lane * K);
    unsigned rs32 = 0u; /* this lane's share of sum_i u(i,B), packed per half */
    NwbPkRange<R> rg;
#pragma unroll
    for (int r = 0; r < R; r++) {
        rg.gcnt[r] = (B - 1 - r >= 0) ? (unsigned)((B - 1 - r) / R + 1) : 0u;
    }
    rg.capg = (B - 1) / R;
    rg.rB = (B - 1) % R;
    rg.ngroups = (unsigned)ngroups;
    /* columns of my 16-byte flush chunk that are inside the table (bit 0 of each nibble) */
    uint4 colmask;
    {
        const int sub_ = lane % ((32 * K) / 16);
        unsigned mm[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            int hi = A - (c * W + sub_ * 32 + q * 8); /* valid nibbles of this word */
            hi = hi < 0 ? 0 : (hi > 8 ? 8 : hi);
            mm[q] = (hi >= 8) ? 0x11111111u : (0x11111111u & ((1u << (4 * hi)) - 1u));
        }
        colmask.x = mm[0]; colmask.y = mm[1]; colmask.z = mm[2]; colmask.w = mm[3];
    }
    unsigned long long *dbg = p.debug_times ? p.debug_times + 4 * (size_t)c : nullptr;
    if (dbg && lane == 0) dbg[0] = nwb_globaltimer();
    unsigned chars_next[NWB_PK_SUB];
#pragma unroll
    for (int t = 0; t < NWB_PK_SUB; t++) {
        chars_next[t] = nwb_pk_chars<R, SMEMCH>(sp_lane + R * t);
    }

    const int nsteps = ngroups + 63;
    const int nblocks = (nsteps + 31) / 32;
    unsigned long long *trace = nullptr;
    unsigned long long npolls = 0;
    if (p.debug_trace && (c % p.debug_trace_stride) == 0 && c / p.debug_trace_stride < 8)
        trace = p.debug_trace + (size_t)(c / p.debug_trace_stride) * p.debug_trace_blocks * 2;
    for (int blk = 0; blk < nblocks; blk++) {
        const int s0 = 32 * blk;
        if (trace && lane == 0 && blk < p.debug_trace_blocks) {
            trace[2 * blk] = nwb_globaltimer();
            trace[2 * blk + 1] = npolls;
        }
        /* The unchecked step is safe everywhere except where the bottom-row sums are captured:
         * rows above the table are virtual (BIG), rows below it compute garbage that is never
         * flushed, and the stream words of row groups outside [0, ngroups) land in the padding of
         * the boundary buffers.  Only the LAST strip captures (the others' share of r(A,B) is read
         * off the boundary stream afterwards), and only once its first lane reaches row B. */
        const bool lean = !(is_last && R * (s0 + 32) >= B);
#pragma unroll 1
        for (int sub = 0; sub < 32 / NWB_PK_SUB; sub++) {
            const int ss = s0 + NWB_PK_SUB * sub;
            if (has_left) {
                /* commit the prefetched words of groups ss .. ss+7; re-poll the ones not valid yet */
                const int gs = ss + lane;
                unsigned w = bq_next;
                if (!nowait) {
                    bool ok = (lane >= NWB_PK_SUB) || (gs >= ngroups) || (w & NWB_PK_VALID);
                    if (!__all_sync(NWB_FULL_MASK, ok)) {
                        /* plain spin, no sleep quantum and no bookkeeping inside the inner loop: its wake-up latency
                         * sits on the strip-to-strip critical path (see nwb_fill_hx.cuh) */
                        NwbWatchdog wd;
                        for (;;) {
                            bool arrived = false;
#pragma unroll 1
                            for (int it = 0; it < NWB_WD_POLLS; it++) {
                                if (!ok) {
                                    w = nwb_ld_relaxed_u32(in_w + gs, left_remote);
                                    ok = (w & NWB_PK_VALID) != 0u;
                                }
#ifdef NWB_EMU
                                nwb_pause();
#endif
                                if (__all_sync(NWB_FULL_MASK, ok)) {
                                    arrived = true;
                                    break;
                                }
                            }
                            npolls += NWB_WD_POLLS;
                            if (arrived) break;
                            if (wd.slow(NWB_ERR_WORD(p), p.watchdog_ns)) return false;
                        }
                    }
                }
                if (dbg && lane == 0 && ss == 0) dbg[1] = nwb_globaltimer();
                if (dbg && lane == 0 && ss == 64) dbg[2] = nwb_globaltimer();
                bq = w & VMASK;
                bq_next = 0u;
                if (lane < NWB_PK_SUB && gs + NWB_PK_SUB < ngroups)
                    bq_next = nwb_ld_relaxed_u32(in_w + gs + NWB_PK_SUB, left_remote);
                if (COUNT) {
                    /* stage the counts of rows R*ss+1 .. R*(ss+SUB): lane i polls word i of the
                     * sub-block (2 words per row), even lanes assemble and store the count */
                    const int nw = NWB_PK_SUB * R * 2;
                    const int row0 = lane >> 1;                 /* row within the sub-block */
                    const bool need = (lane < nw) && (ss + row0 / R < ngroups);
                    unsigned long long cw = 0ull;
                    if (need) cw = nwb_ld_relaxed_u64(in_c + (size_t)ss * R * 2 + lane, left_remote);
                    if (!nowait) {
                        bool okc = !need || (cw & NWB_PK_CVALID);
                        if (!__all_sync(NWB_FULL_MASK, okc)) {
                            NwbWatchdog wdc;
                            for (;;) {
                                bool arrived = false;
#pragma unroll 1
                                for (int it = 0; it < NWB_WD_POLLS; it++) {
                                    if (!okc) {
                                        cw = nwb_ld_relaxed_u64(in_c + (size_t)ss * R * 2 + lane, left_remote);
                                        okc = (cw & NWB_PK_CVALID) != 0ull;
                                    }
#ifdef NWB_EMU
                                    nwb_pause();
#endif
                                    if (__all_sync(NWB_FULL_MASK, okc)) {
                                        arrived = true;
                                        break;
                                    }
                                }
                                if (arrived) break;
                                if (wdc.slow(NWB_ERR_WORD(p), p.watchdog_ns)) return false;
                            }
                        }
                    }
                    const unsigned long long hi = __shfl_down_sync(NWB_FULL_MASK, cw, 1);
                    __syncwarp(); /* the previous sub-block's reads of cstage are done */
                    if (lane < nw && !(lane & 1)) cstage[row0] = (cw & ~NWB_PK_CVALID) | (hi << 63);
                    __syncwarp();
                }
            }
            /* side characters of the NEXT sub-block's steps are loaded one sub-block ahead */
            unsigned chars[NWB_PK_SUB];
#pragma unroll
            for (int t = 0; t < NWB_PK_SUB; t++) chars[t] = nwb_pin_copy(chars_next[t]); /* before the next loads are issued: see nwb_fill_hx.cuh */
            {
                const uint16_t *spn = sp_lane + R * (ss + NWB_PK_SUB);
#pragma unroll
                for (int t = 0; t < NWB_PK_SUB; t++) {
                    chars_next[t] = nwb_pk_chars<R, SMEMCH>(spn + R * t);
                }
            }
            uint32_t *outb = out_w + (ss - 2 * lane - 1);
            unsigned long long *outcb = COUNT ? out_c + (ptrdiff_t)(ss - 2 * lane - 1) * R * 2 : nullptr;
            const int gb = ss - 2 * lane - 1;
#define NWB_PK_SLOT_PTR(t) (lane_stage + (unsigned)(((gb + (t)) & (NWB_PK_RING_ROWS - 1)) * SLOT))
            if (COUNT) {
                /* the count step is ~4x the code of the plain one: keep it rolled so that the loop
                 * body stays inside the instruction cache (side characters come straight from L1) */
                const uint16_t *spc = sp_lane + R * ss;
                if (lean) {
#pragma unroll 1
                    for (int t = 0; t < NWB_PK_SUB; t++)
                        nwb_pk_step<K, R, true, COUNT>(st, pc, rg, bq, t, lane, t, gb + t, A, col_lo, col_hi,
                                                       nwb_pk_chars<R, SMEMCH>(spc + R * t), NWB_PK_SLOT_PTR(t), outb,
                                                       pub31, rs32, cc, cstage, outcb, cfinal);
                } else {
#pragma unroll 1
                    for (int t = 0; t < NWB_PK_SUB; t++)
                        nwb_pk_step<K, R, false, COUNT>(st, pc, rg, bq, t, lane, t, gb + t, A, col_lo, col_hi,
                                                        nwb_pk_chars<R, SMEMCH>(spc + R * t), NWB_PK_SLOT_PTR(t), outb,
                                                        pub31, rs32, cc, cstage, outcb, cfinal);
                }
            } else if (lean) {
#pragma unroll
                for (int t = 0; t < NWB_PK_SUB; t++)
                    nwb_pk_step<K, R, true, COUNT>(st, pc, rg, bq, t, lane, t, gb + t, A, col_lo, col_hi, chars[t],
                                                   NWB_PK_SLOT_PTR(t), outb, pub31, rs32, cc, cstage, outcb, cfinal);
            } else {
#pragma unroll
                for (int t = 0; t < NWB_PK_SUB; t++)
                    nwb_pk_step<K, R, false, COUNT>(st, pc, rg, bq, t, lane, t, gb + t, A, col_lo, col_hi, chars[t],
                                                    NWB_PK_SLOT_PTR(t), outb, pub31, rs32, cc, cstage, outcb, cfinal);
            }
        }
        __syncwarp();
        /* groups <= 32*blk-32 are complete in every lane: move the 32*R newest complete rows from
         * the ring to the arrow table with 16-byte loads/stores (flipping the inverted codes) */
        {
            const int jhi = R * (32 * blk - 31);
            const int jlo = jhi - 32 * R + 1;
            const int row_bytes = 32 * K;          /* bytes per strip row          */
            const int lanes_per_row = row_bytes / 16;
            const int rows_per_pass = 32 / lanes_per_row;
            const int sub = lane % lanes_per_row;
            const int npass = (32 * R) / rows_per_pass;
            /* row j sits at ring offset ((j-1) mod (128*R)) * row_bytes: slots are consecutive row groups */
            const int j0 = jlo + lane / lanes_per_row;
            uint8_t *dst = p.arrows + (size_t)c * row_bytes + sub * 16;
            if (jlo >= 1 && jhi <= B) {
                /* the common case, all rows inside the table: all loads first, then all stores */
#pragma unroll
                for (int half = 0; half < npass; half += 8) {
                    uint4 v[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        if (half + q < npass) {
                            const int j = j0 + (half + q) * rows_per_pass;
                            v[q] = *reinterpret_cast<const uint4 *>(stage_bytes + (size_t)((j - 1) & (NWB_PK_RING_ROWS * R - 1)) * row_bytes + sub * 16);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        if (half + q < npass) {
                            const int j = j0 + (half + q) * rows_per_pass;
                            uint4 w = v[q];
                            w.x = ~w.x & 0x77777777u; w.y = ~w.y & 0x77777777u;
                            w.z = ~w.z & 0x77777777u; w.w = ~w.w & 0x77777777u;
                            *reinterpret_cast<uint4 *>(dst + (size_t)(j - 1) * p.pitch) = w;
                            if (count_branches) branches += nwb_pk_branches(w, colmask);
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int q = 0; q < npass; q++) {
                    const int j = j0 + q * rows_per_pass;
                    if (j >= 1 && j <= B) {
                        uint4 w = *reinterpret_cast<const uint4 *>(stage_bytes + (size_t)((j - 1) & (NWB_PK_RING_ROWS * R - 1)) * row_bytes + sub * 16);
                        w.x = ~w.x & 0x77777777u; w.y = ~w.y & 0x77777777u;
                        w.z = ~w.z & 0x77777777u; w.w = ~w.w & 0x77777777u;
                        *reinterpret_cast<uint4 *>(dst + (size_t)(j - 1) * p.pitch) = w;
                        if (count_branches) branches += nwb_pk_branches(w, colmask);
                    }
                }
            }
        }
        __syncwarp();
    }
    rsum += (long long)(rs32 & 0xFFFFu) + (long long)(rs32 >> 16);
    if (dbg && lane == 0) dbg[3] = nwb_globaltimer();
    return true;
}

/* per-warp shared memory: the arrow ring, then (count kernel) the staged stream counts */
#define NWB_PK_CSTAGE_BYTES (NWB_PK_SUB * 2 * 8)
#define NWB_PK_WARP_SMEM(K, R, COUNT) ((size_t)NWB_PK_RING_ROWS * NWB_PK_SLOT_BYTES(K, R) + ((COUNT) ? NWB_PK_CSTAGE_BYTES : 0))

template <int K, int R, bool COUNT>
__global__ void __launch_bounds__(32 * NWB_PK_MAX_WARPS, 1) nwb_fill_pk_kernel(const NwbStripParams p, const NwbPkConsts pc)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nworkers = (int)gridDim.x * (int)(blockDim.x >> 5);
    const int worker = warp * (int)gridDim.x + (int)blockIdx.x;
    unsigned char *stage = NWB_SMEM_BASE() + (size_t)warp * NWB_PK_WARP_SMEM(K, R, COUNT);
    unsigned long long *cstage = reinterpret_cast<unsigned long long *>(stage + (size_t)NWB_PK_RING_ROWS * NWB_PK_SLOT_BYTES(K, R));

    long long rsum = 0;
    unsigned long long cfinal = 0ull;
    bool owns_final = false;
    unsigned branches = 0;
    for (int c = p.strip_begin + worker; c < p.strip_end; c += nworkers) {
        if (!nwb_pk_strip<K, R, false, COUNT>(p, pc, c, stage, lane, rsum, cstage, cfinal, p.count_branches != 0, branches))
            return; /* watchdog: the results of this launch are invalid (summary->error is set) */
        if (COUNT && c == p.n_strips - 1) {
            /* the block that owns column A wrote cfinal when it passed row B */
            const int W = 64 * K, col_lo = c * W + (2 * lane) * K + 1;
            owns_final = (p.A >= col_lo && p.A < col_lo + 2 * K);
        }
    }
    if (COUNT && owns_final) p.summary->count = cfinal;

#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        rsum += __shfl_xor_sync(NWB_FULL_MASK, rsum, o);
        branches += __shfl_xor_sync(NWB_FULL_MASK, branches, o);
    }
    if (lane == 0 && rsum) atomicAdd((unsigned long long *)&p.summary->rsum, (unsigned long long)rsum);
    if (lane == 0 && branches) atomicAdd(&p.summary->branch_count, branches);
}

/* r(A,B) = [sum over the rows 1..B of v in the column left of the last strip] + [sum of u(i,B) over
 * the last strip's columns].  The second term is captured by the last strip; this adds the first
 * one from the boundary stream that strip consumed (strip n-2's words, all valid by now). */
__global__ void nwb_pk_stream_sum_kernel(const uint32_t *stream, int B, int R, long long *rsum)
{
    const uint32_t *w = stream + NWB_PK_BPAD;
    const int ngroups = (B + R - 1) / R;
    long long s = 0;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += gridDim.x * blockDim.x) {
        const unsigned x = w[g];
        if (R == 1) {
            s += (x >> 16) & 0x7FFFu;
        } else {
            s += x & 0x7FFFu;
            if (2 * g + 2 <= B) s += (x >> 16) & 0x7FFFu;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(NWB_FULL_MASK, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd((unsigned long long *)rsum, (unsigned long long)s);
}

/* ---- branch counter over the finished nibble table -----------------------------
 * get_branch_count() (walk-table.c:133-147): number of interior cells with two
 * or more arrows.  The packed kernel does not count in its inner loop; this
 * HBM-bound pass (0.5 B/cell) produces the counter when it is asked for. */
__global__ void nwb_branch_count_kernel(const uint8_t *arrows, size_t pitch, int A, int B, int col_begin, int col_end,
                                        unsigned *out)
{
    /* columns [col_begin, col_end) (0-based interior), in whole 32-cell groups of 16 B; a block walks
     * rows grid-stride, its threads walk the row's groups: coalesced 16-byte loads, no divisions */
    const int g0 = col_begin / 32, g1 = (col_end + 31) / 32;
    const int cend = col_end < A ? col_end : A;
    unsigned cnt = 0;
    for (int row = blockIdx.x; row < B; row += gridDim.x) {
        const uint8_t *rp = arrows + (size_t)row * pitch;
        for (int g = g0 + (int)threadIdx.x; g < g1; g += (int)blockDim.x) {
            const uint4 v = *reinterpret_cast<const uint4 *>(rp + (size_t)g * 16);
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
            const int first = g * 32;
            const bool edge = (first < col_begin) || (first + 32 > cend);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const unsigned x = w[q];
                const unsigned b0 = x & 0x11111111u, b1 = (x >> 1) & 0x11111111u, b2 = (x >> 2) & 0x11111111u;
                unsigned two = (b0 & b1) | (b0 & b2) | (b1 & b2);
                if (edge) { /* keep columns in [col_begin, cend) */
                    int lo = col_begin - (first + q * 8), hi = cend - (first + q * 8);
                    if (lo < 0) lo = 0;
                    if (hi > 8) hi = 8;
                    unsigned mask = 0u;
                    if (hi > lo) mask = (hi >= 8 ? 0xFFFFFFFFu : ((1u << (4 * hi)) - 1u)) & ~((lo <= 0) ? 0u : ((1u << (4 * lo)) - 1u));
                    two &= mask;
                }
                cnt += (unsigned)__popc(two);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(NWB_FULL_MASK, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out, cnt);
}

#ifndef NWB_EMU
typedef int (*nwb_fail_fn)(cudaError_t, const char *);

template <int K, int R, bool COUNT>
static int nwb_pk_launch_k(const NwbStripParams &sp, const NwbPkConsts &pc, int grid, int warps, cudaStream_t st,
                           nwb_fail_fn fail)
{
    auto kernel = nwb_fill_pk_kernel<K, R, COUNT>;
    const size_t smem = (size_t)warps * NWB_PK_WARP_SMEM(K, R, COUNT);
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    void *args[] = {(void *)&sp, (void *)&pc};
    e = cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(32 * warps), args, smem, st);
    if (e != cudaSuccess) return fail(e, "cudaLaunchCooperativeKernel");
    return 0;
}

static inline int nwb_pk_launch(const NwbStripParams &sp, const NwbPkConsts &pc, int K, int R, bool count, int grid,
                                int warps, cudaStream_t st, nwb_fail_fn fail)
{
    if (count) {
        if (K != 4) return -1;
        return R == 2 ? nwb_pk_launch_k<4, 2, true>(sp, pc, grid, warps, st, fail)
                      : nwb_pk_launch_k<4, 1, true>(sp, pc, grid, warps, st, fail);
    }
    switch (K * 10 + R) {
    case 11: return nwb_pk_launch_k<1, 1, false>(sp, pc, grid, warps, st, fail);
    case 21: return nwb_pk_launch_k<2, 1, false>(sp, pc, grid, warps, st, fail);
    case 41: return nwb_pk_launch_k<4, 1, false>(sp, pc, grid, warps, st, fail);
    case 12: return nwb_pk_launch_k<1, 2, false>(sp, pc, grid, warps, st, fail);
    case 22: return nwb_pk_launch_k<2, 2, false>(sp, pc, grid, warps, st, fail);
    case 42: return nwb_pk_launch_k<4, 2, false>(sp, pc, grid, warps, st, fail);
    default: return -1;
    }
}
#endif
