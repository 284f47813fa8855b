/*
 * nwb_fill_hx.cuh -- packed 16x2 difference fill with flush warps ("hx").
 *
 * Same recurrence, strip pipeline and boundary streams as nwb_fill_pk.cuh (and
 * therefore the same results as the reference's score_cell(),
 * needleman-wunsch.c:418-510), for K = 4 columns per half-lane and R = 2 rows per
 * step, when every per-cell difference fits a nibble (2d + m <= 7).
 *
 * Why.  A strip is swept by ONE warp, and one warp issues at most one
 * instruction every ~2 cycles (measured on B200: tools/ubench/tput.cu), so the
 * strip-to-strip critical path is the sweeping warp's instruction count.  In
 * nwb_fill_pk.cuh a step of 16 cells per lane is ~134 instructions, 60 of them
 * spent on turning differences into arrow nibbles and on the write-back.  Here the
 * sweeping ("critical") warp only computes the recurrence and keeps the three
 * differences of a row's cells as nibbles of two words (SWAR, integer
 * multiply-adds):
 *     P1 = un | 8*[DIAG],   P2 = 8 - vn    (un, vn: the cell's u and v, 0..7)
 * DIAG <=> z - a == 0 comes out of  A4 - Z4 + 0x8888 (bit 3 of each nibble set iff the
 * nibble difference is zero), Z4, A4, NU = the row's z, a, u packed by Horner's rule.
 * Each critical warp streams {P1, P2} through a shared-memory ring to a FLUSH warp
 * on the fourth SM sub-partition, which extracts the zero tests
 *     LEFT <=> un == 0,  UP <=> vn == 0     (bit 3 of 8 - x)
 * builds the reference's arrow sets (needleman-wunsch.c:485-503) as 4-bit codes,
 * counts the branch cells (walk-table.c:108-120) and writes the table, one whole
 * 128-byte strip row per store instruction.
 *
 * Block = 12 warps: warps 0..2 sweep strips (sub-partitions 0..2), warps 3, 7, 11
 * (all on sub-partition 3) flush for warps 0, 1, 2; the other warps exit at once.
 */
#pragma once
#include "nwb_fill_pk.cuh"

#define NWB_HX_CRIT 3
#define NWB_HX_WARPS 12
#define NWB_HX_RING_STEPS 128 /* ring slots, one per step of the sweeping warp */
#define NWB_HX_SLOT_BYTES 512 /* 32 lanes x {P1, P2 of sub-row 0, P1, P2 of sub-row 1} */
#define NWB_HX_RING_BYTES (NWB_HX_RING_STEPS * NWB_HX_SLOT_BYTES)
#define NWB_HX_SMEM_BYTES ((size_t)NWB_HX_CRIT * NWB_HX_RING_BYTES + 64)
#define NWB_HX_B8 0x88888888u
#define NWB_HX_B7 0x77777777u

static inline bool nwb_hx_supported(const NwbPkConsts &pc) { return pc.a_match <= 7; }

struct NwbHxState {
    unsigned tpw[4];   /* pre-shifted top characters of my columns (low block | high block) */
    unsigned u[4];     /* u of my columns in the row above                                   */
    unsigned vlast[2]; /* v of my last columns per sub-row                                   */
    unsigned sp[2];    /* ~(side char << shift) per sub-row (low half row | high half row)   */
    unsigned send;     /* v of my HIGH block's last column for the 2 rows just done          */
    unsigned nu;       /* u of my columns in the row above, one nibble per column            */
};

__device__ __forceinline__ void nwb_sts128(nwb_smem_addr a, unsigned x, unsigned y, unsigned z, unsigned w)
{
#ifdef NWB_EMU
    unsigned *q = reinterpret_cast<unsigned *>(a);
    q[0] = x; q[1] = y; q[2] = z; q[3] = w;
#else
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w));
#endif
}

/* block-scope flags in shared memory between a sweeping warp and its flush warp */
__device__ __forceinline__ int nwb_flag_load(const volatile int *p) { return *p; }
__device__ __forceinline__ void nwb_flag_store(volatile int *p, int v)
{
#if !defined(NWB_EMU) && !defined(NWB_TEST_NOFENCE)
    __threadfence_block();
#endif
    *p = v;
}
__device__ __forceinline__ void nwb_spin_pause(bool sleep)
{
#ifdef NWB_EMU
    (void)sleep;
    nwb_pause();
#else
    if (sleep) __nanosleep(400);
#endif
}

/* One row step of one lane of a sweeping warp: 2 rows x 8 cells (see nwb_pk_step for the
 * lane/half geometry).  slot = this lane's 16 bytes of the step's ring slot. */
template <bool LEAN>
__device__ __forceinline__ void nwb_hx_step(NwbHxState &st, const NwbPkConsts &pc, const NwbPkRange<2> &rg,
                                             const unsigned bq, const int t, const int lane, const int g_idx,
                                             const int g_hi, const int A, const int col_lo, const int col_hi,
                                             const unsigned chars, const nwb_smem_addr slot, uint32_t *out_w,
                                             const bool pub31, unsigned &rs32)
{
    unsigned recv = __shfl_up_sync(NWB_FULL_MASK, st.send, 1);
    const unsigned b = __shfl_sync(NWB_FULL_MASK, bq, t);
    if (lane == 0) recv = b;
    unsigned vL[2];
    vL[0] = __byte_perm(recv, st.vlast[0], 0x5410); /* lo <- neighbour's row 0, hi <- my low block's row 0 */
    vL[1] = __byte_perm(recv, st.vlast[1], 0x5432);
    st.sp[0] = __byte_perm(chars, st.sp[0], 0x5410);
    st.sp[1] = __byte_perm(chars, st.sp[1], 0x5432);
    unsigned uafter[2][4];
    unsigned p1[2], p2[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        unsigned v = vL[r];
        unsigned z[4], a[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned nx = st.tpw[k] ^ st.sp[r];                  /* -x'-1 per half   */
            a[k] = __viaddmax_s16x2(nx, pc.TT1, pc.AMIS);             /* a_match or a_mis */
            z[k] = __vimax3_s16x2(a[k], v, st.u[k]);
            const unsigned un = z[k] - v;
            const unsigned vn = z[k] - st.u[k];
            st.u[k] = un;
            uafter[r][k] = un;
            v = vn;
        }
        st.vlast[r] = v;
        /* the row's z, a and u as one nibble per column (per half: 4 columns) */
        const unsigned Z4 = ((z[3] * 16u + z[2]) * 16u + z[1]) * 16u + z[0];
        const unsigned A4 = ((a[3] * 16u + a[2]) * 16u + a[1]) * 16u + a[0];
        const unsigned NU = ((st.u[3] * 16u + st.u[2]) * 16u + st.u[1]) * 16u + st.u[0];
        const unsigned ZT = A4 - Z4 + NWB_HX_B8; /* bit 3 of a nibble: z == a (DIAG)            */
        const unsigned ZV = st.nu - Z4 + NWB_HX_B8; /* vn = z - u(row above); bit 3: vn == 0 (UP) */
        st.nu = NU;
        p1[r] = (ZT & NWB_HX_B8) | NU;
        p2[r] = ZV;
    }
    nwb_sts128(slot, p1[0], p2[0], p1[1], p2[1]);
    st.send = __byte_perm(st.vlast[0], st.vlast[1], 0x7632);
    /* bottom row, r(A,B) = sum of u(i,B): row B passes through a lane's low block in one step and
     * through its high block in the next (see nwb_pk_step) */
    if (!LEAN && __builtin_expect((unsigned)(rg.capg - g_hi) <= 1u, 0)) {
        const unsigned half = (g_hi == rg.capg) ? 0xFFFF0000u : 0x0000FFFFu;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned x = uafter[0][k];
            if (rg.rB == 1) x = uafter[1][k];
            unsigned m = 0u;
            if (col_lo + k <= A) m |= 0x0000FFFFu;
            if (col_hi + k <= A) m |= 0xFFFF0000u;
            rs32 += x & m & half;
        }
    }
    /* lane 31: the strip's last column for group g_hi, self-validating (nwb_fill_pk.cuh) */
    const bool pub = pub31 && (LEAN || (unsigned)g_hi < rg.ngroups);
    nwb_st_relaxed_sys_pred(out_w + g_idx, st.send | 0x80008000u, pub);
}

/* A sweeping warp's strip: nwb_pk_strip<4, 2> without the write-back.  seq counts this warp's
 * 32-step blocks over all its strips; ring slot of a step = (32 * seq + step in block) mod 128. */
/* Pipelined strip group (nwb_plan_run_pipelined): the strip that streams into the right neighbour's inbox writes
 * into copy e & 1 of it, which carried fill e - 2: wait until the neighbour says it is done with that fill (in steady
 * state it long is).  In the kernel, not as a kernel of its own in front of the fill: a spinning kernel at the head of
 * a stream held up the other plans' streams (measured: the fills of a queue ran in waves).  Not inlined, and called
 * before the warp's strip loop: the sweeping loop's instruction schedule must not depend on it. */
__device__ __noinline__ bool nwb_hx_gate(const uint32_t *ack, const unsigned need, int *err, const unsigned long long limit_ns)
{
    NwbWatchdog wd;
    for (;;) {
        const unsigned v = nwb_ld_relaxed_u32(ack, true);
        if (__all_sync(NWB_FULL_MASK, (int)(v - need) >= 0)) return true;
        if (wd.tick(err, limit_ns)) return false;
#ifdef NWB_EMU
        nwb_pause();
#endif
    }
}

template <bool QUEUE>
__device__ __forceinline__ bool nwb_hx_strip(const NwbStripParams &p, const NwbPkConsts &pc, const int c,
                                              unsigned char *ring, volatile int *ready, volatile int *done,
                                              int &seq, const int lane, long long &rsum)
{
    const int K = 4, R = 2;
    const int A = p.A, B = p.B;
    const int W = 64 * K;
    const int col_lo = c * W + (2 * lane) * K + 1;
    const int col_hi = col_lo + K;
    const unsigned ONE = 0x00010001u;
    const int ngroups = (B + R - 1) / R;

    NwbHxState st;
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
    }
    st.send = NWB_PK_BIG * ONE;
    st.nu = 0u;

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
    const uint16_t *sp_lane = p.side_pre + NWB_PK_SPAD + 1 - 2 * R * lane;
    const unsigned VMASK = 0x7FFF7FFFu;
    unsigned bq = 0u, bq_next = 0u;
    if (has_left && lane < NWB_PK_SUB && lane < ngroups) bq_next = nwb_ld_relaxed_u32(in_w + lane, left_remote);

    const nwb_smem_addr lane_ring = nwb_smem_address(ring) + (unsigned)(lane * 16);
    unsigned rs32 = 0u;
    NwbPkRange<R> rg;
#pragma unroll
    for (int r = 0; r < R; r++) rg.gcnt[r] = (B - 1 - r >= 0) ? (unsigned)((B - 1 - r) / R + 1) : 0u;
    rg.capg = (B - 1) / R;
    rg.rB = (B - 1) % R;
    rg.ngroups = (unsigned)ngroups;

    unsigned chars_next[NWB_PK_SUB];
#pragma unroll
    for (int t = 0; t < NWB_PK_SUB; t++) chars_next[t] = nwb_pk_chars<R, false>(sp_lane + R * t);

    const int nsteps = ngroups + 63;
    const int nblocks = (nsteps + 31) / 32;
    for (int blk = 0; blk < nblocks; blk++) {
        const int s0 = 32 * blk;
        /* ring back-pressure: this block overwrites the slots of block seq-4; the rows the flush
         * warp takes after block seq-2 still read them */
        if (seq >= 2 && nwb_flag_load(done) < seq - 1) {
            /* The flush warp only ever waits for THIS warp, so it always gets there -- unless it left because some
             * watchdog fired: look at the error word now and then.  (A full watchdog tick in this loop changed the
             * sweeping warp's schedule and cost 33 % of the whole fill.) */
            for (;;) {
                bool freed = false;
#pragma unroll 1
                for (int it = 0; it < NWB_WD_POLLS; it++) {
                    if (nwb_flag_load(done) >= seq - 1) {
                        freed = true;
                        break;
                    }
#ifdef NWB_EMU
                    nwb_pause();
#endif
                }
                if (freed) break;
                const int *ew = NWB_ERR_WORD(p);
                if (ew && *reinterpret_cast<const volatile int *>(ew) != 0) return false;
            }
        }
        const bool lean = !(is_last && R * (s0 + 32) >= B);
#pragma unroll 1
        for (int sub = 0; sub < 32 / NWB_PK_SUB; sub++) {
            const int ss = s0 + NWB_PK_SUB * sub;
            if (has_left) {
                /* commit the prefetched words of groups ss .. ss+7; re-poll the ones not valid yet */
                const int gs = ss + lane;
                unsigned w = bq_next;
                bool ok = (lane >= NWB_PK_SUB) || (gs >= ngroups) || (w & NWB_PK_VALID) || NWB_DBG_BITS(p, 1);
                if (!__all_sync(NWB_FULL_MASK, ok)) {
                    /* The wake-up latency of this loop sits on the strip-to-strip critical path (391 hops at 100k):
                     * the inner loop is load, vote, leave; the watchdog's bookkeeping runs outside it, once per
                     * NWB_WD_POLLS polls (inside the loop it cost 33 % of the whole fill). */
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
                        if (arrived) break;
                        if (wd.slow(NWB_ERR_WORD(p), p.watchdog_ns)) return false;
                    }
                }
                bq = w & VMASK;
                bq_next = 0u;
                if (lane < NWB_PK_SUB && gs + NWB_PK_SUB < ngroups)
                    bq_next = nwb_ld_relaxed_u32(in_w + gs + NWB_PK_SUB, left_remote);
            }
            /* The side characters of this sub-block were loaded one sub-block ago; they are COPIED out of their load
             * registers here, before the next sub-block's loads are issued.  Global loads retire through a counting
             * scoreboard: a step that read a load register directly would also wait for the younger loads issued
             * just below -- an L2 round trip per sub-block, +33 % on the whole fill (measured when ptxas renamed the
             * registers instead of copying).  The volatile move pins the copy and its place. */
            unsigned chars[NWB_PK_SUB];
#pragma unroll
            for (int t = 0; t < NWB_PK_SUB; t++) chars[t] = nwb_pin_copy(chars_next[t]);
            {
                const uint16_t *spn = sp_lane + R * (ss + NWB_PK_SUB);
#pragma unroll
                for (int t = 0; t < NWB_PK_SUB; t++) chars_next[t] = nwb_pk_chars<R, false>(spn + R * t);
            }
            uint32_t *outb = out_w + (ss - 2 * lane - 1);
            const int gb = ss - 2 * lane - 1;
            const nwb_smem_addr slot0 =
                lane_ring + (unsigned)(((32 * seq + NWB_PK_SUB * sub) & (NWB_HX_RING_STEPS - 1)) * NWB_HX_SLOT_BYTES);
            if (lean) {
#pragma unroll
                for (int t = 0; t < NWB_PK_SUB; t++)
                    nwb_hx_step<true>(st, pc, rg, bq, t, lane, t, gb + t, A, col_lo, col_hi, chars[t],
                                      slot0 + (unsigned)(t * NWB_HX_SLOT_BYTES), outb, pub31, rs32);
            } else {
#pragma unroll
                for (int t = 0; t < NWB_PK_SUB; t++)
                    nwb_hx_step<false>(st, pc, rg, bq, t, lane, t, gb + t, A, col_lo, col_hi, chars[t],
                                       slot0 + (unsigned)(t * NWB_HX_SLOT_BYTES), outb, pub31, rs32);
            }
        }
        __syncwarp();
        if (lane == 0) nwb_flag_store(ready, seq + 1);
        seq++;
    }
    rsum += (long long)(rs32 & 0xFFFFu) + (long long)(rs32 >> 16);
    return true;
}

/* The flush warp of sweeping warp `wslot`.  Flush lane h owns the 8 cells per row that sweeping
 * lane h computes (4 bytes of every arrow row of the strip; the warp's 32 lanes store one whole
 * 128-byte strip row per instruction).  Sweeping lane h left, at step s, the words of {low half:
 * its low block in group s - 2h, high half: its high block in group s - 2h - 1}; both halves of
 * group g are therefore in the slots of steps g + 2h and g + 2h + 1, and walking the groups in
 * order reads every slot once.  After block blk of a strip the groups 32*blk-63 .. 32*blk-32 are
 * complete in every lane. */
/* Arrow codes of the 8 cells of one ring word pair {p1, p2} (both halves; the caller knows which
 * half belongs to which row group).  zv = 0x8888.. - vn was formed by the sweeping warp. */
__device__ __forceinline__ unsigned nwb_hx_code(const unsigned p1, const unsigned zv, unsigned &two)
{
    const unsigned zu = NWB_HX_B8 - (p1 & NWB_HX_B7);                             /* bit 3: un == 0 (LEFT) */
    two = (p1 & zu) | (p1 & zv) | (zu & zv);                                      /* bit 3: two or more arrows */
    const unsigned q = (((zu >> 1) & 0x44444444u) | (zv & NWB_HX_B8)) >> 1;       /* LEFT at bit 1, UP at bit 2 */
    return ((p1 >> 3) & 0x11111111u) | (q & 0x66666666u);
}

/* One ring slot of my lane: codes of {low half: group gl, high half: group gl - 1}, both sub-rows.
 * Rows of group g are complete once the slots of steps g + 2h and g + 2h + 1 are both seen: the
 * low half comes from the previous slot's codes (pc0/pc1), the high half from this one. */
template <bool COUNT, bool EDGE>
__device__ __forceinline__ void nwb_hx_flush_slot(const uint4 w, unsigned &pc0, unsigned &pc1, uint8_t *out,
                                                  const size_t pitch, const bool row1, const unsigned colmask,
                                                  const unsigned halfmask, unsigned &branches)
{
    unsigned t0, t1;
    const unsigned c0 = nwb_hx_code(w.x, w.y, t0);
    const unsigned c1 = nwb_hx_code(w.z, w.w, t1);
    if (COUNT) {
        /* both halves of a slot are cells of the table (different row groups); EDGE: first / last
         * slots of a strip, where one half lies above row 1 or below row B */
        const unsigned m0 = EDGE ? (colmask & halfmask) : colmask;
        branches += (unsigned)__popc(t0 & m0);
        if (!EDGE || row1) branches += (unsigned)__popc(t1 & m0);
    }
    *reinterpret_cast<unsigned *>(out) = __byte_perm(pc0, c0, 0x7610);
    if (!EDGE || row1) *reinterpret_cast<unsigned *>(out + pitch) = __byte_perm(pc1, c1, 0x7610);
    pc0 = c0;
    pc1 = c1;
}

/* worker / nworkers: the first strip (relative to strip_begin) and the strip stride of the sweeping warp this warp
 * flushes for */
template <bool PUBLISH>
__device__ __forceinline__ void nwb_hx_flush(const NwbStripParams &p, const int worker, const int nworkers,
                                              const unsigned char *ring, volatile int *ready, volatile int *done,
                                              const int lane, unsigned &branches)
{
    const int B = p.B, A = p.A;
    const int ngroups = (B + 1) / 2;
    const int nfull = B / 2; /* groups with both rows inside the table */
    const int nblocks = (ngroups + 63 + 31) / 32;
    const bool count_branches = p.count_branches != 0;
    const unsigned char *lane_ring = ring + lane * 16;
    int seq = 0;
    for (int c = p.strip_begin + worker; c < p.strip_end; c += nworkers) {
        /* bit 3 of the nibbles of my word whose column is inside the table */
        int hi = A - (c * 256 + lane * 8);
        hi = hi < 0 ? 0 : (hi > 8 ? 8 : hi);
        const unsigned colmask = (hi >= 8) ? NWB_HX_B8 : (NWB_HX_B8 & ((1u << (4 * hi)) - 1u));
        uint8_t *dst = p.arrows + (size_t)c * 128 + (size_t)lane * 4;
        unsigned pc0 = 0u, pc1 = 0u;
        for (int blk = 0; blk < nblocks; blk++) {
            {
                NwbWatchdog wd;
                while (nwb_flag_load(ready) < seq + 1) {
                    nwb_spin_pause(true);
                    if (wd.tick(NWB_ERR_WORD(p), p.watchdog_ns)) return; /* the sweeping warp gave up (or never came) */
                }
            }
#ifndef NWB_EMU
            __threadfence_block();
#endif
            if (!NWB_DBG_BITS(p, 2)) {
                /* slot of step s holds {low: group s - 2*lane, high: group s - 2*lane - 1}; walk the slots
                 * whose HIGH half is one of this block's groups 32*blk-63 .. 32*blk-32 (clipped to the
                 * table); the slot before the first one only primes pc0/pc1 */
                const int sbase = 32 * (seq - blk) + 2 * lane + 1; /* ring step of the slot whose high half is group 0 */
                int g = 32 * blk - 63;
                int gend = g + 32;
                if (g < 0) g = 0;
                if (gend > ngroups) gend = ngroups;
                if (g < gend) {
                    const unsigned char *q = lane_ring;
                    if (g == 0) { /* the strip's first group: its low halves sit in the slot before */
                        const uint4 w = *reinterpret_cast<const uint4 *>(q + ((sbase - 1) & (NWB_HX_RING_STEPS - 1)) * NWB_HX_SLOT_BYTES);
                        unsigned t;
                        pc0 = nwb_hx_code(w.x, w.y, t);
                        if (count_branches) branches += (unsigned)__popc(t & colmask & 0x0000FFFFu);
                        pc1 = nwb_hx_code(w.z, w.w, t);
                        if (count_branches && B >= 2) branches += (unsigned)__popc(t & colmask & 0x0000FFFFu);
                    }
                    uint8_t *out = dst + (size_t)(2 * g) * p.pitch;
                    const int gfull = gend < nfull ? gend : nfull;
                    /* a slot's low half belongs to group g + 1: not a table cell for the last group,
                     * and only its first sub-row when B is odd and g + 1 is the last group */
                    const int gbulk = gfull < ngroups - 2 ? gfull : (ngroups - 2 < g ? g : ngroups - 2);
#pragma unroll 4
                    for (; g < gbulk; g++) {
                        const uint4 w = *reinterpret_cast<const uint4 *>(q + ((sbase + g) & (NWB_HX_RING_STEPS - 1)) * NWB_HX_SLOT_BYTES);
                        if (count_branches) nwb_hx_flush_slot<true, false>(w, pc0, pc1, out, p.pitch, true, colmask, 0u, branches);
                        else nwb_hx_flush_slot<false, false>(w, pc0, pc1, out, p.pitch, true, colmask, 0u, branches);
                        out += 2 * p.pitch;
                    }
                    for (; g < gend; g++) { /* the last two groups of the strip: per-half, per-row validity */
                        const uint4 w = *reinterpret_cast<const uint4 *>(q + ((sbase + g) & (NWB_HX_RING_STEPS - 1)) * NWB_HX_SLOT_BYTES);
                        const bool row1 = 2 * g + 2 <= B;
                        unsigned t0, t1;
                        const unsigned c0 = nwb_hx_code(w.x, w.y, t0);
                        const unsigned c1 = nwb_hx_code(w.z, w.w, t1);
                        if (count_branches) {
                            /* high half: group g (rows 2g+1, 2g+2); low half: group g+1 (rows 2g+3, 2g+4) */
                            unsigned m0 = colmask & 0xFFFF0000u, m1 = row1 ? (colmask & 0xFFFF0000u) : 0u;
                            if (2 * g + 3 <= B) m0 |= colmask & 0x0000FFFFu;
                            if (2 * g + 4 <= B) m1 |= colmask & 0x0000FFFFu;
                            branches += (unsigned)__popc(t0 & m0) + (unsigned)__popc(t1 & m1);
                        }
                        *reinterpret_cast<unsigned *>(out) = __byte_perm(pc0, c0, 0x7610);
                        if (row1) *reinterpret_cast<unsigned *>(out + p.pitch) = __byte_perm(pc1, c1, 0x7610);
                        pc0 = c0;
                        pc1 = c1;
                        out += 2 * p.pitch;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) nwb_flag_store(done, seq + 1);
            if (PUBLISH) {
                /* rows 1 .. 2 * (last group of this block) of the strip are written: tell the count sweep */
                int gdone = 32 * blk - 31;
                gdone = gdone < 0 ? 0 : (gdone > ngroups ? ngroups : gdone);
                const int rows = 2 * gdone < B ? 2 * gdone : B;
#ifndef NWB_EMU
                __threadfence(); /* my stores before the flag */
#endif
                __syncwarp();
                if (lane == 0 && rows > 0) nwb_st_relaxed_u32(reinterpret_cast<uint32_t *>(p.progress + (c - p.strip_begin)), (unsigned)rows, false);
            }
            seq++;
        }
    }
}

/* PUBLISH: the flush warps also publish in p.progress[] how many arrow rows of each strip are in memory, for a
 * count sweep (nwb_count.cuh) that trails the fill on another stream.  A separate instantiation: the sweeping
 * warps' instruction schedule is sensitive to any code added to the kernel. */
template <bool PUBLISH, bool QUEUE = false>
__global__ void __launch_bounds__(32 * NWB_HX_WARPS, 1) nwb_fill_hx_kernel(const NwbStripParams p, const NwbPkConsts pc)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *smem = NWB_SMEM_BASE();
    volatile int *flags = reinterpret_cast<volatile int *>(smem + (size_t)NWB_HX_CRIT * NWB_HX_RING_BYTES);
    if (threadIdx.x < 2 * NWB_HX_CRIT) flags[threadIdx.x] = 0;
    /* Two ways of dealing the strips out (p.hx_spb):
     * 0: strips b, b + G, b + 2G for block b of a grid of G <= SMs blocks that are all resident (cooperative launch) --
     *    the three sweeping warps of an SM are far apart in the table and the first G strips have an SM each to
     *    themselves: the shortest time for ONE fill.
     * 1..3 (queue mode): block ids are TICKETS drawn as the blocks start and a block sweeps that many ADJACENT strips.
     *    A strip then only ever waits for strips of blocks with smaller tickets, i.e. blocks that are already running
     *    or done, so the launch needs no co-residency guarantee, a table may have more blocks than the GPU has SMs,
     *    and the blocks of the NEXT fill (another plan, another stream) move onto SMs as this fill's blocks leave
     *    them: consecutive fills overlap. */
    /* QUEUE is a template parameter: the instantiation for one fill alone is the code it was before queue mode
     * existed (the sweeping warp's schedule is sensitive to anything added to the kernel: +2 % with a run-time switch) */
    const bool queue = QUEUE;
    if (QUEUE && threadIdx.x == 0) flags[2 * NWB_HX_CRIT] = p.summary ? atomicAdd(&p.summary->ticket, 1) : (int)blockIdx.x;
    __syncthreads();
    const int bid = QUEUE ? flags[2 * NWB_HX_CRIT] : (int)blockIdx.x;
    const int spb = (QUEUE && p.hx_spb >= 1 && p.hx_spb <= NWB_HX_CRIT) ? p.hx_spb : NWB_HX_CRIT; /* sweeping warps of this block */
    const int nworkers = (int)gridDim.x * spb;
    const int crit_slot = (warp < spb) ? warp : -1;
    const int flush_slot = ((warp & 3) == 3 && (warp >> 2) < spb) ? (warp >> 2) : -1;
    if (crit_slot >= 0) {
        /* sweeping warp */
        const int worker = queue ? bid * spb + crit_slot : crit_slot * (int)gridDim.x + bid;
        unsigned char *ring = smem + (size_t)crit_slot * NWB_HX_RING_BYTES;
        long long rsum = 0;
        int seq = 0;
        if (QUEUE && p.gate_ack && p.strip_end < p.n_strips) {
            /* am I the warp that sweeps this launch's last strip? */
            const int last = p.strip_end - 1 - p.strip_begin;
            if (last >= worker && (last - worker) % nworkers == 0 &&
                !nwb_hx_gate(p.gate_ack, p.gate_need, NWB_ERR_WORD(p), p.watchdog_ns))
                return;
        }
        for (int c = p.strip_begin + worker; c < p.strip_end; c += nworkers)
            if (!nwb_hx_strip<QUEUE>(p, pc, c, ring, flags + crit_slot, flags + NWB_HX_CRIT + crit_slot, seq, lane, rsum))
                return; /* watchdog: summary->error is set, the host reports NWB_ERR_CUDA */
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(NWB_FULL_MASK, rsum, o);
        if (lane == 0 && rsum) atomicAdd((unsigned long long *)&p.summary->rsum, (unsigned long long)rsum);
    } else if (flush_slot >= 0) {
        /* flush warp of sweeping warp flush_slot */
        const int wslot = flush_slot;
        unsigned branches = 0;
        nwb_hx_flush<PUBLISH>(p, queue ? bid * spb + wslot : wslot * (int)gridDim.x + bid, nworkers, smem + (size_t)wslot * NWB_HX_RING_BYTES, flags + wslot,
                              flags + NWB_HX_CRIT + wslot, lane, branches);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) branches += __shfl_xor_sync(NWB_FULL_MASK, branches, o);
        if (lane == 0 && branches) atomicAdd(&p.summary->branch_count, branches);
    }
}

#ifndef NWB_EMU
template <bool PUBLISH, bool QUEUE>
static int nwb_hx_launch_t(const NwbStripParams &sp, const NwbPkConsts &pc, int grid, cudaStream_t st, nwb_fail_fn fail)
{
    auto kernel = nwb_fill_hx_kernel<PUBLISH, QUEUE>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NWB_HX_SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    if (QUEUE) {
        /* an ordinary launch: the blocks' tickets make it deadlock-free without co-residency (see the kernel) */
        kernel<<<dim3(grid), dim3(32 * NWB_HX_WARPS), NWB_HX_SMEM_BYTES, st>>>(sp, pc);
        e = cudaGetLastError();
        if (e != cudaSuccess) return fail(e, "nwb_fill_hx_kernel launch");
        return 0;
    }
    void *args[] = {(void *)&sp, (void *)&pc};
    e = cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(32 * NWB_HX_WARPS), args, NWB_HX_SMEM_BYTES, st);
    if (e != cudaSuccess) return fail(e, "cudaLaunchCooperativeKernel");
    return 0;
}
static inline int nwb_hx_launch(const NwbStripParams &sp, const NwbPkConsts &pc, int grid, cudaStream_t st, nwb_fail_fn fail)
{
    if (sp.hx_spb > 0) /* queue mode; the count sweep that trails the fill (publish_rows) belongs to the one-fill mode */
        return nwb_hx_launch_t<false, true>(sp, pc, grid, st, fail);
    return sp.publish_rows ? nwb_hx_launch_t<true, false>(sp, pc, grid, st, fail) : nwb_hx_launch_t<false, false>(sp, pc, grid, st, fail);
}
#endif
