/*
 * nwb_count.cuh -- 64-bit path count over the finished arrow table.
 *
 * The number of optimal alignments behind `-s` (get_solution_count(),
 * computation.c:249-260; the reference obtains it by enumerating every alignment,
 * needleman-wunsch.c:209-331) is the number of arrow paths from (A,B) to (0,0):
 *     cnt(0,0) = cnt(i,0) = cnt(0,j) = 1,
 *     cnt(i,j) = [DIAG] cnt(i-1,j-1) + [LEFT] cnt(i-1,j) + [UP] cnt(i,j-1)      (mod 2^64)
 * whose low 32 bits are what the reference's `unsigned int solution_count` holds.
 *
 * nwb_fill_pk.cuh can fuse this recurrence into the fill (three 64-bit multiply-adds
 * per cell in the sweeping warp: the step grows from ~134 to ~600 instructions and
 * the fill from 10 to 44 ms at 100k x 100k).  This kernel instead runs it as a
 * second sweep over the 4-bit arrow codes the fill has written (0.5 B/cell re-read
 * through L1/L2), with the same strip pipeline but its own geometry: one warp per
 * strip of 32*CPL columns, lane l owns CPL columns and works on row s - l at step s
 * (32 steps of skew per strip), the left neighbour's count arrives by two shuffles,
 * strips hand their last column's counts on as two self-validating 64-bit words per
 * row (bit 63 = valid; word 0 = bits 0..62, word 1 = bit 63), over peer memory when
 * the next strip lives on another GPU.  Per cell: six selects and one three-input
 * add with carry (IADD3 + IADD3.X), all on the ALU pipe, so the sweep is bound by one
 * warp's issue rate on that pipe (~22 cycles per cell in isolation, tools/ubench/cnt.cu).
 * CPL = 8, 4 or 2 cells per lane and row (nwb_count_choose_cpl).
 */
#pragma once
#include "nwb_fill_pk.cuh"
#include "nwb_digest.cuh"
#include "nwb_count_sparse.cuh"

#define NWB_CNT_WARPS 4   /* one per SM sub-partition */
#define NWB_CNT_SUB 8     /* rows per sub-block: arrow words and stream words are fetched one sub-block ahead */
#define NWB_CNT_RING 64   /* arrow-word ring slots per warp (>= 31 steps of lane skew + 3 sub-blocks) */
#define NWB_CNT_SMEM_PER_WARP (NWB_CNT_SUB * 8 + NWB_CNT_RING * 32 * 4)
#define NWB_CNT_SMEM_BYTES (NWB_CNT_WARPS * NWB_CNT_SMEM_PER_WARP)

struct NwbCountParams {
    const uint8_t *arrows; /* nibble table, B rows x pitch bytes (include/nwb.h layout) */
    size_t pitch;
    int A, B;
    int n_strips;    /* strips of 32*CPL columns over the whole table    */
    int strip_begin; /* this launch (GPU) sweeps [strip_begin, strip_end) */
    int strip_end;
    unsigned long long *bnd_c; /* [strip - strip_begin][2 * bpitch] count streams of the local strips */
    size_t bpitch;
    const unsigned long long *in_bnd_c; /* inbox written by the left-neighbour GPU (strip_begin > 0)   */
    unsigned long long *out_bnd_c;      /* right neighbour's inbox (strip_end < n_strips)              */
    NwbDevSummary *summary;
    const int *fill_progress; /* CPL == 8 only: [strip - strip_begin] arrow rows of the strip the fill has written so
                               * far, when the sweep runs concurrently with nwb_fill_hx_kernel; NULL = table is finished */
    const int *skip_state; /* if non-NULL and *skip_state == NWB_SPC_DONE the kernel returns at once: the sparse
                            * backward sweep (nwb_count_sparse.cuh) already produced the count */
    int debug_nowait; /* diagnostics: 1 = do not wait for the left strip's stream (results are wrong) */
    unsigned long long watchdog_ns; /* bounded waits: see NwbStripParams */
};

template <int CPL>
struct NwbCntWord;
template <>
struct NwbCntWord<8> { typedef uint32_t T; };
template <>
struct NwbCntWord<4> { typedef uint16_t T; };
template <>
struct NwbCntWord<2> { typedef uint8_t T; };

template <typename T>
__device__ __forceinline__ unsigned nwb_cnt_load(const T *p)
{
#ifdef NWB_EMU
    return (unsigned)*p;
#else
    return (unsigned)__ldca(p); /* cached in L1: the 32 lanes of a warp read one row's bytes over 32 steps */
#endif
}

__device__ __forceinline__ uint4 nwb_ldg_u128(const uint8_t *p)
{
#ifdef NWB_EMU
    uint4 v;
    memcpy(&v, p, 16);
    return v;
#else
    return __ldcg(reinterpret_cast<const uint4 *>(p)); /* streamed once: L2 only */
#endif
}

/* One row of one lane: CPL cells.  x = the lane's CPL arrow nibbles of the row. */
template <int CPL>
__device__ __forceinline__ void nwb_count_row(const unsigned x, unsigned long long (&cnt)[CPL], unsigned long long &left_above,
                                               unsigned long long cl, unsigned long long &send)
{
    unsigned long long cd = left_above;
    left_above = cl;
#pragma unroll
    for (int k = 0; k < CPL; k++) {
        const unsigned f = x >> (4 * k);
        const unsigned long long cu = cnt[k];
        const unsigned long long n = ((f & 1u) ? cd : 0ull) + ((f & 2u) ? cl : 0ull) + ((f & 4u) ? cu : 0ull);
        cd = cu;
        cnt[k] = n;
        cl = n;
    }
    send = cl;
}

/* One strip: columns c*32*CPL+1 .. (c+1)*32*CPL, all rows. */
template <int CPL, bool DIGEST>
__device__ __forceinline__ bool nwb_count_strip(const NwbCountParams &p, const int c, unsigned long long *cstage,
                                                 unsigned *ring, const int lane, unsigned long long &dig_row,
                                                 unsigned long long &dig_col)
{
    typedef typename NwbCntWord<CPL>::T word_t;
    const int A = p.A, B = p.B;
    const int W = 32 * CPL;
    const int lc = c - p.strip_begin;
    const bool has_left = (c > 0);
    const bool left_remote = has_left && (lc == 0);
    const bool publish = (c + 1 < p.n_strips);
    const bool out_remote = publish && (c == p.strip_end - 1);
    /* stream of row j (1-based): words 2*(BPAD + j - 1), +1 */
    unsigned long long *out_c = (out_remote ? p.out_bnd_c : p.bnd_c + (size_t)lc * 2 * p.bpitch) + (size_t)NWB_PK_BPAD * 2;
    const unsigned long long *in_c = nullptr;
    if (has_left)
        in_c = (left_remote ? p.in_bnd_c : p.bnd_c + (size_t)(lc - 1) * 2 * p.bpitch) + (size_t)NWB_PK_BPAD * 2;
    const bool pub31 = publish && (lane == 31) && !NWB_FAULT_INJECTED(p);
    bool aborted = false;
    /* the cell (A, B): the lane / column that owns it */
    const int kfin = (c == p.n_strips - 1) ? (A - 1 - c * W) - CPL * lane : -1; /* 0..CPL-1 in the owning lane */

    unsigned long long cnt[CPL];
#pragma unroll
    for (int k = 0; k < CPL; k++) cnt[k] = 1ull; /* border row */
    unsigned long long send = 1ull;               /* my last column in the row I just finished (border: 1) */
    unsigned long long left_above = 1ull;         /* count left of my first column in the row above        */
    const word_t *wp = reinterpret_cast<const word_t *>(p.arrows + (size_t)c * (W / 2) + (size_t)lane * (CPL / 2));
    const size_t wpitch = p.pitch / sizeof(word_t);

    if (!has_left && lane < NWB_CNT_SUB) cstage[lane] = 1ull; /* column 0 of the table */
    __syncwarp();

    /* CPL == 8: the arrow rows go through a shared-memory ring.  The warp loads 8 rows x 128 bytes per
     * sub-block with two 16-byte loads per lane (coalesced, two sub-blocks ahead of use) and scatters the
     * words so that lane l's word of row r sits in slot (r + l - 1) mod 64: at step s every lane reads
     * slot s mod 64, an immediate offset inside an unrolled sub-block.  (Loading each lane's own word
     * straight from the table touches 32 different rows per instruction.) */
    const bool staged = (CPL == 8);
    uint4 stage0 = make_uint4(0u, 0u, 0u, 0u), stage1 = stage0;
    const int srow = lane >> 3, schunk = lane & 7; /* my row (0..3, +4) and 16-byte chunk of a staged block */
    auto stage_load = [&](const int rbase) { /* rows rbase .. rbase+7 into stage0/1 */
        const uint8_t *q = p.arrows + (size_t)c * (W / 2) + (size_t)schunk * 16;
        const int r0 = rbase + srow, r1 = r0 + 4;
        stage0 = (r0 >= 1 && r0 <= B) ? nwb_ldg_u128(q + (size_t)(r0 - 1) * p.pitch) : make_uint4(0u, 0u, 0u, 0u);
        stage1 = (r1 >= 1 && r1 <= B) ? nwb_ldg_u128(q + (size_t)(r1 - 1) * p.pitch) : make_uint4(0u, 0u, 0u, 0u);
    };
    auto stage_store = [&](const int rbase) {
        const int r0 = rbase + srow, r1 = r0 + 4, w0 = 4 * schunk;
        const unsigned v0[4] = {stage0.x, stage0.y, stage0.z, stage0.w}, v1[4] = {stage1.x, stage1.y, stage1.z, stage1.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            ring[((r0 + w0 + i - 1) & (NWB_CNT_RING - 1)) * 32 + w0 + i] = v0[i];
            ring[((r1 + w0 + i - 1) & (NWB_CNT_RING - 1)) * 32 + w0 + i] = v1[i];
        }
    };
    /* arrow words of my rows of the next sub-block (direct path), stream words of lane 0's rows of the next
     * two sub-blocks */
    unsigned wnext[NWB_CNT_SUB];
    const int *fillp = (staged && p.fill_progress) ? p.fill_progress + lc : nullptr;
    int rows_seen = 0; /* the fill publishes 64 rows at a time: poll (and fence) only when the cached value is short */
    auto wait_rows = [&](const int need) {
        if (rows_seen >= need || aborted) return;
        NwbWatchdog wd;
        do {
            rows_seen = (int)nwb_ld_relaxed_u32(reinterpret_cast<const uint32_t *>(fillp), false);
            if (rows_seen < need && wd.tick(NWB_ERR_WORD(p), p.watchdog_ns)) {
                aborted = true; /* the fill on the other stream is not making progress (not co-resident?) */
                return;
            }
#ifdef NWB_EMU
            if (rows_seen < need) nwb_pause();
#endif
        } while (rows_seen < need);
#ifndef NWB_EMU
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
    };
    if (staged) {
        __syncwarp(); /* the previous strip's reads of the ring are done */
        if (fillp) {
            rows_seen = 0;
            wait_rows(B < 2 * NWB_CNT_SUB ? B : 2 * NWB_CNT_SUB);
            if (aborted) return false;
        }
        stage_load(1);
        stage_store(1);
        __syncwarp();
        stage_load(1 + NWB_CNT_SUB);
    } else {
#pragma unroll
        for (int t = 0; t < NWB_CNT_SUB; t++) {
            const int j = t - lane + 1;
            wnext[t] = (j >= 1 && j <= B) ? nwb_cnt_load(wp + (size_t)(j - 1) * wpitch) : 0u;
        }
    }
    unsigned long long cw_next = 0ull, cw_next2 = 0ull;
    if (has_left && lane < 2 * NWB_CNT_SUB) {
        if ((lane >> 1) < B) cw_next = nwb_ld_relaxed_u64(in_c + lane, left_remote);
        if ((lane >> 1) + NWB_CNT_SUB < B) cw_next2 = nwb_ld_relaxed_u64(in_c + 2 * NWB_CNT_SUB + lane, left_remote);
    }

    const int nsteps = B + 31;
    /* One sub-block of 8 steps.  cwslot holds the stream words of this sub-block (fetched two sub-blocks ago) and
     * takes the words of the sub-block after next: the two slots alternate (the caller's loop is unrolled by two),
     * so the fresh load lands in a register nobody reads for two sub-blocks.  With one rotating pair of variables
     * the compiler keeps the load in a temporary and copies it at the end of the prologue, and that copy waits for
     * the L2 round trip in every sub-block: 22 % of the sweep's time (ncu source view). */
    auto sub_block = [&](const int ss, unsigned long long &cwslot) {
        if (has_left) {
            /* lane 0's left inputs for rows ss+1 .. ss+8: 16 stream words, lane i takes word i */
            const int row = ss + (lane >> 1) + 1;
            const bool need = (lane < 2 * NWB_CNT_SUB) && (row <= B);
            unsigned long long cw = cwslot;
            bool ok = !need || (cw & NWB_PK_CVALID) || NWB_DBG_BITS(p, 1);
            if (!__all_sync(NWB_FULL_MASK, ok)) {
                NwbWatchdog wd; /* its bookkeeping stays outside the inner poll loop (see nwb_fill_hx.cuh) */
                for (;;) {
                    bool arrived = false;
#pragma unroll 1
                    for (int it = 0; it < NWB_WD_POLLS; it++) {
                        if (!ok) {
                            cw = nwb_ld_relaxed_u64(in_c + (size_t)ss * 2 + lane, left_remote);
                            ok = (cw & NWB_PK_CVALID) != 0ull;
                        }
#ifdef NWB_EMU
                        nwb_pause();
#endif
                        if (__all_sync(NWB_FULL_MASK, ok)) {
                            arrived = true;
                            break;
                        }
                    }
                    if (arrived) break;
                    if (wd.slow(NWB_ERR_WORD(p), p.watchdog_ns)) {
                        aborted = true;
                        return;
                    }
                }
            }
            const unsigned long long hi = __shfl_down_sync(NWB_FULL_MASK, cw, 1);
            __syncwarp(); /* the previous sub-block's reads of cstage are done */
            if (lane < 2 * NWB_CNT_SUB && !(lane & 1)) cstage[lane >> 1] = (cw & ~NWB_PK_CVALID) | (hi << 63);
            __syncwarp();
        }
        unsigned w[NWB_CNT_SUB];
        if (staged) {
            /* rows ss+9 .. ss+16 (loaded during the previous sub-block) into the ring; rows ss+17 .. ss+24 on their way */
            stage_store(ss + 1 + NWB_CNT_SUB);
            if (fillp) { /* rows up to ss + 24 are about to be read from the table */
                const int need = ss + 3 * NWB_CNT_SUB;
                wait_rows(need < B ? need : B);
                if (aborted) return;
            }
            const unsigned *rq = ring + (ss & (NWB_CNT_RING - 1)) * 32 + lane;
#pragma unroll
            for (int t = 0; t < NWB_CNT_SUB; t++) w[t] = rq[t * 32];
            __syncwarp(); /* this sub-block's ring reads are issued before the next sub-block's stores */
            /* every global load of the sub-block is issued here, AFTER the last use of the previous sub-block's
             * loads (stage_store above, cw at the top): loads retire through a counting scoreboard, so a use that
             * follows a younger load waits for that one too -- an L2 round trip per sub-block, 22 % of the sweep
             * (ncu source view), when the stream-word prefetch was issued before stage_store */
            stage_load(ss + 1 + 2 * NWB_CNT_SUB);
        } else {
#pragma unroll
            for (int t = 0; t < NWB_CNT_SUB; t++) w[t] = wnext[t];
        }
        /* stream words are fetched two sub-blocks ahead (a word that is not valid yet costs an L2 round trip);
         * unconditional: rows beyond B land in the stream's padding and are never looked at */
        if (has_left) cwslot = nwb_ld_relaxed_u64(in_c + (size_t)(ss + 2 * NWB_CNT_SUB) * 2 + lane, left_remote);
        /* every lane strictly inside rows 1 .. B-1 for this sub-block and the next one's loads inside the table */
        const bool lean = (ss >= 31) && (ss + 2 * NWB_CNT_SUB < B);
        if (lean) {
            if (!staged) {
                const word_t *wq = wp + (size_t)(ss + NWB_CNT_SUB - lane) * wpitch;
#pragma unroll
                for (int t = 0; t < NWB_CNT_SUB; t++) wnext[t] = nwb_cnt_load(wq + (size_t)t * wpitch);
            }
            unsigned long long *oc = out_c + (size_t)(ss - lane) * 2;
#pragma unroll
            for (int t = 0; t < NWB_CNT_SUB; t++) {
                unsigned long long cl = __shfl_up_sync(NWB_FULL_MASK, send, 1);
                if (lane == 0) cl = cstage[t];
                nwb_count_row<CPL>(w[t], cnt, left_above, cl, send);
                if (DIGEST && kfin >= 0 && kfin < CPL) { /* the lane that owns column A: cnt(A, j) */
#pragma unroll
                    for (int k = 0; k < CPL; k++)
                        if (k == kfin) dig_col += nwb_mix64((unsigned long long)(ss + t - lane + 1), cnt[k]);
                }
                nwb_st_relaxed_sys_pred_u64(oc + 2 * t, send | NWB_PK_CVALID, pub31);
                nwb_st_relaxed_sys_pred_u64(oc + 2 * t + 1, (send >> 63) | NWB_PK_CVALID, pub31);
            }
        } else {
            if (!staged) {
#pragma unroll
                for (int t = 0; t < NWB_CNT_SUB; t++) {
                    const int j = ss + NWB_CNT_SUB + t - lane + 1;
                    wnext[t] = (j >= 1 && j <= B) ? nwb_cnt_load(wp + (size_t)(j - 1) * wpitch) : 0u;
                }
            }
#pragma unroll 1
            for (int t = 0; t < NWB_CNT_SUB; t++) {
                const int j = ss + t - lane + 1; /* my row at this step */
                unsigned long long cl = __shfl_up_sync(NWB_FULL_MASK, send, 1);
                if (lane == 0) cl = cstage[t];
                if (j >= 1 && j <= B) {
                    unsigned x = w[0];
#pragma unroll
                    for (int q = 1; q < NWB_CNT_SUB; q++)
                        if (q == t) x = w[q];
                    nwb_count_row<CPL>(x, cnt, left_above, cl, send);
                    if (DIGEST) {
                        if (kfin >= 0 && kfin < CPL) {
#pragma unroll
                            for (int k = 0; k < CPL; k++)
                                if (k == kfin) dig_col += nwb_mix64((unsigned long long)j, cnt[k]);
                        }
                        if (j == B) { /* bottom row: cnt(i, B) of my columns inside the table */
#pragma unroll
                            for (int k = 0; k < CPL; k++) {
                                const int col = c * W + CPL * lane + k + 1;
                                if (col <= A) dig_row += nwb_mix64((unsigned long long)col, cnt[k]);
                            }
                        }
                    }
                    if (j == B && kfin >= 0 && kfin < CPL) {
#pragma unroll
                        for (int k = 0; k < CPL; k++)
                            if (k == kfin) p.summary->count = cnt[k];
                    }
                    nwb_st_relaxed_sys_pred_u64(out_c + (size_t)(j - 1) * 2, send | NWB_PK_CVALID, pub31);
                    nwb_st_relaxed_sys_pred_u64(out_c + (size_t)(j - 1) * 2 + 1, (send >> 63) | NWB_PK_CVALID, pub31);
                }
            }
        }
    };
    for (int ss = 0; ss < nsteps && !aborted; ss += 2 * NWB_CNT_SUB) {
        sub_block(ss, cw_next);
        if (ss + NWB_CNT_SUB < nsteps && !aborted) sub_block(ss + NWB_CNT_SUB, cw_next2);
    }
    return !aborted;
}

/* DIGEST: also accumulate the digests of the bottom row and of column A of the count matrix
 * (NwbDevSummary.dig_row / dig_col, nwb_digest.cuh) -- a separate instantiation for parity checks. */
template <int CPL, bool DIGEST>
__global__ void __launch_bounds__(32 * NWB_CNT_WARPS, 1) nwb_count_kernel(const NwbCountParams p)
{
    if (p.skip_state && *reinterpret_cast<const volatile int *>(p.skip_state) == NWB_SPC_DONE) return;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nworkers = (int)gridDim.x * NWB_CNT_WARPS;
    const int worker = warp * (int)gridDim.x + (int)blockIdx.x;
    unsigned long long *cstage = reinterpret_cast<unsigned long long *>(NWB_SMEM_BASE() + (size_t)warp * NWB_CNT_SMEM_PER_WARP);
    unsigned *ring = reinterpret_cast<unsigned *>(cstage + NWB_CNT_SUB);
    unsigned long long dig_row = 0ull, dig_col = 0ull;
    for (int c = p.strip_begin + worker; c < p.strip_end; c += nworkers)
        if (!nwb_count_strip<CPL, DIGEST>(p, c, cstage, ring, lane, dig_row, dig_col)) return; /* watchdog */
    if (DIGEST) {
        dig_row = nwb_warp_sum_u64(dig_row);
        dig_col = nwb_warp_sum_u64(dig_col);
        if (lane == 0 && dig_row) atomicAdd(&p.summary->dig_row, dig_row);
        if (lane == 0 && dig_col) atomicAdd(&p.summary->dig_col, dig_col);
    }
}

/* Cells per lane and row.  Narrower strips shorten a row step but add a pipeline hop of ~45 steps per
 * strip; measured on B200 (10k, 30k, 100k squares) 8 cells per lane (256-column strips) is the fastest
 * or within 3 % of it everywhere: one step costs ~200 cycles of shuffle, stream and load overhead plus
 * ~16 cycles per cell, so the overhead dominates the narrow variants. */
static inline int nwb_count_choose_cpl(long long columns, int sm_count)
{
    (void)columns;
    (void)sm_count;
    return 8;
}

#ifndef NWB_EMU
template <int CPL, bool DIGEST>
static int nwb_count_launch_t(const NwbCountParams &cp, int grid, cudaStream_t st, nwb_fail_fn fail)
{
    void *args[] = {(void *)&cp};
    /* cooperative launch only for its co-residency guarantee (strips spin-wait on one another) */
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)nwb_count_kernel<CPL, DIGEST>, dim3(grid),
                                                dim3(32 * NWB_CNT_WARPS), args, NWB_CNT_SMEM_BYTES, st);
    if (e != cudaSuccess) return fail(e, "cudaLaunchCooperativeKernel");
    return 0;
}
static inline int nwb_count_launch(const NwbCountParams &cp, int cpl, bool digest, int grid, cudaStream_t st, nwb_fail_fn fail)
{
    if (digest) return nwb_count_launch_t<8, true>(cp, grid, st, fail);
    switch (cpl) {
    case 2: return nwb_count_launch_t<2, false>(cp, grid, st, fail);
    case 4: return nwb_count_launch_t<4, false>(cp, grid, st, fail);
    default: return nwb_count_launch_t<8, false>(cp, grid, st, fail);
    }
}
#endif
