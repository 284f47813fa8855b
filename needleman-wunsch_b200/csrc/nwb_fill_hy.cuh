/*
 * nwb_fill_hy.cuh -- nwb_fill_hx.cuh with THREE rows of skew per lane instead of four ("hy").
 *
 * Same recurrence (score_cell(), needleman-wunsch.c:418-510), same block (three sweeping
 * warps + three flush warps), same ring words {P1, P2}, same self-validating boundary
 * stream and therefore the same results as nwb_fill_hx.cuh / nwb_fill_pk.cuh, whose strips
 * (and peer GPUs) it interoperates with.  What changes is the lane geometry.  In hx a
 * virtual lane (one 16-bit half of a register: 4 columns) runs one whole step, R = 2 rows,
 * behind its left neighbour, so every strip boundary costs 64 steps of skew + 8 of
 * look-ahead on the strip-to-strip critical path (7.3 us per hop, 2.85 of the 7.5 ms at
 * 100k x 100k).  Here a lane's high half runs ONE ROW behind its low half (its left inputs
 * are the lane's own registers: sub-row 0 takes the low half's sub-row 1 of the previous
 * step, sub-row 1 the low half's sub-row 0 of THIS step), and the next lane's low half one
 * step behind that, as in hx (one shuffle per step, at the step boundary):
 *
 *     step s, lane l, rho = 2s - 3l:   low  half: rows rho+1 (sub-row 0), rho+2 (sub-row 1)
 *                                      high half: rows rho   (sub-row 0), rho+1 (sub-row 1)
 *
 * A strip is 48 steps deep instead of 64.
 *
 * STATUS: an experiment kept for reproducibility, NOT the default (NWB_PK_HY=1 selects it;
 * tools/ab_hy.py measures it against hx).  Bit-exact (emulator tests, GPU goldens), same
 * instruction count per step as hx (78), but 4-6 % slower on B200: 7.82 vs 7.53 ms at
 * 100k x 100k, 2.32 vs 2.18 ms at 30k, 0.786 vs 0.741 ms at 10k.  In hx the two rows of a
 * step form a 2 x 4 wavefront (two independent dependency chains that ptxas interleaves,
 * 176 cycles per step = the warp's issue rate); here sub-row 1 cannot start before sub-row
 * 0's last cell, the step is one chain of 8 cells + shuffle, and it takes ~214 cycles, which
 * more than eats the 16 steps saved per hop.  One row of skew per virtual lane (32 steps per
 * strip, a second shuffle in the middle of the step) was measured too: 223 cycles per step,
 * 7.44 ms at 100k but 0.752 / 2.22 ms at 10k / 30k.
 *
 * Stream word of row group g (rows 2g+1, 2g+2): lane 31's high half is on exactly these
 * rows at step g + 47.  Side characters: rho is odd in odd lanes, so a lane reads aligned
 * 32-bit words of side_pre and picks its four rows from two consecutive words with
 * per-lane byte selectors.  Flush lane h finds, in the slot of step s: {sub-row 0: low half
 * of row rho+1, high half of row rho; sub-row 1: low half of row rho+2, high half of row
 * rho+1}; it walks its slots in order and writes rows rho (low half from the previous slot)
 * and rho+1 (both halves from this slot): even lanes rows 2e, 2e+1, odd lanes 2e+1, 2e+2.
 */
#pragma once
#include "nwb_fill_hx.cuh"

struct NwbHyState {
    unsigned tpw[4];   /* pre-shifted top characters of my columns (low block | high block) */
    unsigned u[4];     /* u of my columns in the row above                                   */
    unsigned vlast[2]; /* v of my last columns per sub-row (low block | high block)          */
    unsigned send;     /* v of my HIGH block's last column for the 2 rows just done          */
    unsigned cprev;    /* the previous step's side characters (sub-row 0 | sub-row 1)        */
    unsigned nu;       /* u of my columns in the row above, one nibble per column            */
};

/* One row of 8 cells of one lane: the recurrence, then the row's z, a, u as nibbles. */
__device__ __forceinline__ void nwb_hy_row(NwbHyState &st, const NwbPkConsts &pc, const unsigned sp, unsigned v,
                                            unsigned &vout, unsigned (&uafter)[4], unsigned &p1, unsigned &p2)
{
    unsigned z[4], a[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const unsigned nx = st.tpw[k] ^ sp;                   /* -x'-1 per half   */
        a[k] = __viaddmax_s16x2(nx, pc.TT1, pc.AMIS);        /* a_match or a_mis */
        z[k] = __vimax3_s16x2(a[k], v, st.u[k]);
        const unsigned un = z[k] - v;
        const unsigned vn = z[k] - st.u[k];
        st.u[k] = un;
        uafter[k] = un;
        v = vn;
    }
    vout = v;
    const unsigned Z4 = ((z[3] * 16u + z[2]) * 16u + z[1]) * 16u + z[0];
    const unsigned A4 = ((a[3] * 16u + a[2]) * 16u + a[1]) * 16u + a[0];
    const unsigned NU = ((st.u[3] * 16u + st.u[2]) * 16u + st.u[1]) * 16u + st.u[0];
    const unsigned ZT = A4 - Z4 + NWB_HX_B8;    /* bit 3 of a nibble: z == a (DIAG)            */
    const unsigned ZV = st.nu - Z4 + NWB_HX_B8; /* vn = z - u(row above); bit 3: vn == 0 (UP) */
    st.nu = NU;
    p1 = (ZT & NWB_HX_B8) | NU;
    p2 = ZV;
}

/* One step of one lane of a sweeping warp.  rho = 2s - 3*lane; selA / selB = the lane's byte selectors for the side
 * characters of sub-row 0 / 1 out of {previous word, this word}. */
template <bool LEAN>
__device__ __forceinline__ void nwb_hy_step(NwbHyState &st, const NwbPkConsts &pc, const unsigned ngroups, const int B,
                                             const unsigned bq, const int t, const int lane, const unsigned selA,
                                             const unsigned selB, const int g_idx, const int gpub, const int rho,
                                             const int A, const int col_lo, const int col_hi, const unsigned chars,
                                             const nwb_smem_addr slot, uint32_t *out_w, const bool pub31, unsigned &rs32)
{
    unsigned recv = __shfl_up_sync(NWB_FULL_MASK, st.send, 1);
    const unsigned b = __shfl_sync(NWB_FULL_MASK, bq, t);
    if (lane == 0) recv = b;
    const unsigned vL0 = __byte_perm(recv, st.vlast[1], 0x5410);   /* lo <- neighbour's row, hi <- my low block, previous step */
    const unsigned sp0 = __byte_perm(st.cprev, chars, selA);       /* low: row rho+1, high: row rho   */
    const unsigned sp1 = __byte_perm(st.cprev, chars, selB);       /* low: row rho+2, high: row rho+1 */
    st.cprev = chars;
    unsigned ua0[4], ua1[4], p1a, p2a, p1b, p2b;
    nwb_hy_row(st, pc, sp0, vL0, st.vlast[0], ua0, p1a, p2a);
    const unsigned vL1 = __byte_perm(recv, st.vlast[0], 0x5432);   /* hi <- my low block, sub-row 0 of this step */
    nwb_hy_row(st, pc, sp1, vL1, st.vlast[1], ua1, p1b, p2b);
    nwb_sts128(slot, p1a, p2a, p1b, p2b);
    st.send = __byte_perm(st.vlast[0], st.vlast[1], 0x7632);
    /* bottom row, r(A,B) = sum of u(i,B) (last strip only) */
    if (!LEAN && __builtin_expect((unsigned)(B - rho) <= 2u, 0)) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned mlo = (col_lo + k <= A) ? 0x0000FFFFu : 0u;
            const unsigned mhi = (col_hi + k <= A) ? 0xFFFF0000u : 0u;
            if (rho + 1 == B) rs32 += (ua0[k] & mlo) + (ua1[k] & mhi);
            if (rho + 2 == B) rs32 += ua1[k] & mlo;
            if (rho == B) rs32 += ua0[k] & mhi;
        }
    }
    /* lane 31: the strip's last column for group s - 47, self-validating (nwb_fill_pk.cuh) */
    const bool pub = pub31 && (LEAN || (unsigned)gpub < ngroups);
    nwb_st_relaxed_sys_pred(out_w + g_idx, st.send | 0x80008000u, pub);
}

/* A sweeping warp's strip.  seq counts this warp's 32-step blocks over all its strips; ring slot of a step =
 * (32 * seq + step in block) mod 128. */
__device__ __forceinline__ void nwb_hy_strip(const NwbStripParams &p, const NwbPkConsts &pc, const int c,
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

    NwbHyState st;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const unsigned lo = (col_lo + k <= A) ? (unsigned)p.top[col_lo + k - 1] : 0u;
        const unsigned hi = (col_hi + k <= A) ? (unsigned)p.top[col_hi + k - 1] : 0u;
        st.tpw[k] = ((lo << pc.shift) | ((hi << pc.shift) << 16));
        st.u[k] = 0u;
    }
    st.vlast[0] = NWB_PK_BIG * ONE;
    st.vlast[1] = NWB_PK_BIG * ONE;
    st.send = NWB_PK_BIG * ONE;
    st.cprev = 0xFFFFFFFFu;
    st.nu = 0u;

    const int lc = c - p.strip_begin;
    const bool has_left = (c > 0);
    const bool left_remote = has_left && (lc == 0);
    const bool publish = (c + 1 < p.n_strips);
    const bool out_remote = publish && (c == p.strip_end - 1);
    uint32_t *out_w = (out_remote ? p.out_bnd_w : p.bnd_w + (size_t)lc * p.bpitch) + NWB_PK_BPAD;
    const uint32_t *in_w = nullptr;
    if (has_left) in_w = (left_remote ? p.in_bnd_w : p.bnd_w + (size_t)(lc - 1) * p.bpitch) + NWB_PK_BPAD;
    const bool pub31 = publish && (lane == 31) && !(p.debug_nowait & 4);
    const bool is_last = (c == p.n_strips - 1);
    /* rho = 2s - 3*lane.  Even lanes: this step's aligned word of side_pre = rows (rho+1, rho+2), the previous one
     * (rho-1, rho); odd lanes: this step's word = rows (rho+2, rho+3), the previous one (rho, rho+1). */
    const uint16_t *sp_lane = p.side_pre + NWB_PK_SPAD + 1 - 3 * lane + (lane & 1);
    const unsigned selA = (lane & 1) ? 0x1032u : 0x3254u; /* {low: row rho+1, high: row rho}   out of {previous, this} */
    const unsigned selB = (lane & 1) ? 0x3254u : 0x5476u; /* {low: row rho+2, high: row rho+1}                         */

    const unsigned VMASK = 0x7FFF7FFFu;
    unsigned bq = 0u, bq_next = 0u;
    if (has_left && lane < NWB_PK_SUB && lane < ngroups) bq_next = nwb_ld_relaxed_u32(in_w + lane, left_remote);

    const nwb_smem_addr lane_ring = nwb_smem_address(ring) + (unsigned)(lane * 16);
    unsigned rs32 = 0u;

    unsigned chars_next[NWB_PK_SUB];
#pragma unroll
    for (int t = 0; t < NWB_PK_SUB; t++) chars_next[t] = nwb_pk_chars<R, false>(sp_lane + R * t);

    const int nsteps = ngroups + 48;
    const int nblocks = (nsteps + 31) / 32;
    for (int blk = 0; blk < nblocks; blk++) {
        const int s0 = 32 * blk;
        /* ring back-pressure: this block overwrites the slots of block seq-4; the flush of block n reads the
         * slots of blocks n-2 .. n, so the flush of block seq-2 must be complete */
        if (seq >= 2) {
            while (nwb_flag_load(done) < seq - 1) nwb_spin_pause(false);
        }
        const bool lean = !(is_last && R * (s0 + 32) >= B);
#pragma unroll 1
        for (int sub = 0; sub < 32 / NWB_PK_SUB; sub++) {
            const int ss = s0 + NWB_PK_SUB * sub;
            if (has_left) {
                /* commit the prefetched words of groups ss .. ss+7; re-poll the ones not valid yet */
                const int gs = ss + lane;
                unsigned w = bq_next;
                bool ok = (lane >= NWB_PK_SUB) || (gs >= ngroups) || (w & NWB_PK_VALID) || (p.debug_nowait & 1);
                while (!__all_sync(NWB_FULL_MASK, ok)) {
                    if (!ok) {
                        w = nwb_ld_relaxed_u32(in_w + gs, left_remote);
                        ok = (w & NWB_PK_VALID) != 0u;
                    }
#ifdef NWB_EMU
                    nwb_pause();
#endif
                }
                bq = w & VMASK;
                bq_next = 0u;
                if (lane < NWB_PK_SUB && gs + NWB_PK_SUB < ngroups)
                    bq_next = nwb_ld_relaxed_u32(in_w + gs + NWB_PK_SUB, left_remote);
            }
            unsigned chars[NWB_PK_SUB];
#pragma unroll
            for (int t = 0; t < NWB_PK_SUB; t++) chars[t] = nwb_pin_copy(chars_next[t]); /* before the next loads are issued: see nwb_fill_hx.cuh */
            {
                const uint16_t *spn = sp_lane + R * (ss + NWB_PK_SUB);
#pragma unroll
                for (int t = 0; t < NWB_PK_SUB; t++) chars_next[t] = nwb_pk_chars<R, false>(spn + R * t);
            }
            uint32_t *outb = out_w + (ss - 47);
            const int rb = 2 * ss - 3 * lane;
            const nwb_smem_addr slot0 =
                lane_ring + (unsigned)(((32 * seq + NWB_PK_SUB * sub) & (NWB_HX_RING_STEPS - 1)) * NWB_HX_SLOT_BYTES);
            if (lean) {
#pragma unroll
                for (int t = 0; t < NWB_PK_SUB; t++)
                    nwb_hy_step<true>(st, pc, (unsigned)ngroups, B, bq, t, lane, selA, selB, t, ss + t - 47, rb + 2 * t, A, col_lo, col_hi, chars[t],
                                      slot0 + (unsigned)(t * NWB_HX_SLOT_BYTES), outb, pub31, rs32);
            } else {
#pragma unroll
                for (int t = 0; t < NWB_PK_SUB; t++)
                    nwb_hy_step<false>(st, pc, (unsigned)ngroups, B, bq, t, lane, selA, selB, t, ss + t - 47, rb + 2 * t, A, col_lo, col_hi, chars[t],
                                       slot0 + (unsigned)(t * NWB_HX_SLOT_BYTES), outb, pub31, rs32);
            }
        }
        __syncwarp();
        if (lane == 0) nwb_flag_store(ready, seq + 1);
        seq++;
    }
    rsum += (long long)(rs32 & 0xFFFFu) + (long long)(rs32 >> 16);
}

/* One ring slot of flush lane h: rows rho and rho+1 of the lane's 8 columns (out points at row rho).  pc1 = the codes
 * of the previous slot's sub-row 1 (its low half is row rho). */
template <bool COUNT>
__device__ __forceinline__ void nwb_hy_flush_slot(const uint4 w, unsigned &pc1, uint8_t *out, const size_t pitch,
                                                  const unsigned colmask, unsigned &branches)
{
    unsigned t0, t1;
    const unsigned c0 = nwb_hx_code(w.x, w.y, t0);
    const unsigned c1 = nwb_hx_code(w.z, w.w, t1);
    if (COUNT) branches += (unsigned)__popc(t0 & colmask) + (unsigned)__popc(t1 & colmask);
    *reinterpret_cast<unsigned *>(out) = __byte_perm(pc1, c0, 0x7610);
    *reinterpret_cast<unsigned *>(out + pitch) = __byte_perm(c0, c1, 0x7610);
    pc1 = c1;
}

/* The flush warp of sweeping warp `wslot`.  Lane h's slot of step s covers rho = 2s - 3h; with o = ceil(3h/2) and
 * e = s - o that is rho = 2e in even lanes and 2e + 1 in odd ones.  The lanes walk e together (even lanes write rows
 * 2e, 2e+1, odd lanes 2e+1, 2e+2: two half-rows of 16 x 4 bytes per store instruction, merged in L2).  After block
 * blk of the sweeping warp (steps up to 32*blk + 31) every lane has its slots up to e = 32*blk - 16. */
template <bool PUBLISH>
__device__ __forceinline__ void nwb_hy_flush(const NwbStripParams &p, const int wslot, const unsigned char *ring,
                                              volatile int *ready, volatile int *done, const int lane,
                                              unsigned &branches)
{
    const int B = p.B, A = p.A;
    const int nworkers = (int)gridDim.x * NWB_HX_CRIT;
    const int worker = wslot * (int)gridDim.x + (int)blockIdx.x;
    const int ngroups = (B + 1) / 2;
    const int par = lane & 1;
    const int off = (3 * lane + 1) >> 1;
    const int emax = B / 2;                       /* last e with a row of the table in some lane's slot      */
    const int ebulk = B >= 3 ? (B - 3) / 2 : -1;  /* last e with rows rho .. rho+2 inside the table in every lane */
    const int nblocks = (ngroups + 48 + 31) / 32;
    const bool count_branches = p.count_branches != 0;
    const unsigned char *lane_ring = ring + lane * 16;
    int seq = 0;
    for (int c = p.strip_begin + worker; c < p.strip_end; c += nworkers) {
        /* bit 3 of the nibbles of my word whose column is inside the table */
        int hi = A - (c * 256 + lane * 8);
        hi = hi < 0 ? 0 : (hi > 8 ? 8 : hi);
        const unsigned colmask = (hi >= 8) ? NWB_HX_B8 : (NWB_HX_B8 & ((1u << (4 * hi)) - 1u));
        uint8_t *dst = p.arrows + (size_t)c * 128 + (size_t)lane * 4;
        unsigned pc1 = 0u;
        for (int blk = 0; blk < nblocks; blk++) {
            while (nwb_flag_load(ready) < seq + 1) nwb_spin_pause(true);
#ifndef NWB_EMU
            __threadfence_block();
#endif
            if (!(p.debug_nowait & 2)) {
                const int sbase = 32 * (seq - blk) + off; /* ring step of my slot of e = 0 */
                int e = 32 * blk - 47;
                int eend = 32 * blk - 15;
                if (e < -1) e = -1; /* e = -1: odd lanes' slot that holds the low half of row 1 */
                if (eend > emax + 1) eend = emax + 1;
                const unsigned char *q = lane_ring;
                /* the strip's first slots and its last ones: per-half, per-row validity */
                auto edge = [&](const int ee) {
                    const int rho = 2 * ee + par;
                    if (ee + off < 0) return; /* lane 0 has no slot before step 0 */
                    const uint4 w = *reinterpret_cast<const uint4 *>(q + ((sbase + ee) & (NWB_HX_RING_STEPS - 1)) * NWB_HX_SLOT_BYTES);
                    unsigned t0, t1;
                    const unsigned c0 = nwb_hx_code(w.x, w.y, t0);
                    const unsigned c1 = nwb_hx_code(w.z, w.w, t1);
                    const bool r0 = rho >= 1 && rho <= B, r1 = rho + 1 >= 1 && rho + 1 <= B, r2 = rho + 2 >= 1 && rho + 2 <= B;
                    if (count_branches) {
                        /* sub-row 0: {low: row rho+1, high: row rho}; sub-row 1: {low: row rho+2, high: row rho+1} */
                        const unsigned m0 = (r1 ? 0x0000FFFFu : 0u) | (r0 ? 0xFFFF0000u : 0u);
                        const unsigned m1 = (r2 ? 0x0000FFFFu : 0u) | (r1 ? 0xFFFF0000u : 0u);
                        branches += (unsigned)__popc(t0 & colmask & m0) + (unsigned)__popc(t1 & colmask & m1);
                    }
                    if (r0) *reinterpret_cast<unsigned *>(dst + (size_t)(rho - 1) * p.pitch) = __byte_perm(pc1, c0, 0x7610);
                    if (r1) *reinterpret_cast<unsigned *>(dst + (size_t)rho * p.pitch) = __byte_perm(c0, c1, 0x7610);
                    pc1 = c1;
                };
                for (; e < eend && e < 1; e++) edge(e);
                const int eb = eend < ebulk + 1 ? eend : ebulk + 1;
                if (e < eb) {
                    uint8_t *out = dst + (size_t)(2 * e + par - 1) * p.pitch; /* row rho */
#pragma unroll 4
                    for (; e < eb; e++) {
                        const uint4 w = *reinterpret_cast<const uint4 *>(q + ((sbase + e) & (NWB_HX_RING_STEPS - 1)) * NWB_HX_SLOT_BYTES);
                        if (count_branches) nwb_hy_flush_slot<true>(w, pc1, out, p.pitch, colmask, branches);
                        else nwb_hy_flush_slot<false>(w, pc1, out, p.pitch, colmask, branches);
                        out += 2 * p.pitch;
                    }
                }
                for (; e < eend; e++) edge(e);
            }
            __syncwarp();
            if (lane == 0) nwb_flag_store(done, seq + 1);
            if (PUBLISH) {
                /* every lane has written its rows up to 2 * (32 * blk - 16) + 1: tell the count sweep */
                int rows = 64 * blk - 31;
                rows = rows < B ? rows : B;
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

template <bool PUBLISH>
__global__ void __launch_bounds__(32 * NWB_HX_WARPS, 1) nwb_fill_hy_kernel(const NwbStripParams p, const NwbPkConsts pc)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *smem = NWB_SMEM_BASE();
    volatile int *flags = reinterpret_cast<volatile int *>(smem + (size_t)NWB_HX_CRIT * NWB_HX_RING_BYTES);
    if (threadIdx.x < 2 * NWB_HX_CRIT) flags[threadIdx.x] = 0;
    __syncthreads();
    const int crit_slot = (warp < NWB_HX_CRIT) ? warp : -1;
    const int flush_slot = ((warp & 3) == 3 && (warp >> 2) < NWB_HX_CRIT) ? (warp >> 2) : -1;
    if (crit_slot >= 0) {
        const int nworkers = (int)gridDim.x * NWB_HX_CRIT;
        const int worker = crit_slot * (int)gridDim.x + (int)blockIdx.x;
        unsigned char *ring = smem + (size_t)crit_slot * NWB_HX_RING_BYTES;
        long long rsum = 0;
        int seq = 0;
        for (int c = p.strip_begin + worker; c < p.strip_end; c += nworkers)
            nwb_hy_strip(p, pc, c, ring, flags + crit_slot, flags + NWB_HX_CRIT + crit_slot, seq, lane, rsum);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(NWB_FULL_MASK, rsum, o);
        if (lane == 0 && rsum) atomicAdd((unsigned long long *)&p.summary->rsum, (unsigned long long)rsum);
    } else if (flush_slot >= 0) {
        const int wslot = flush_slot;
        unsigned branches = 0;
        nwb_hy_flush<PUBLISH>(p, wslot, smem + (size_t)wslot * NWB_HX_RING_BYTES, flags + wslot,
                              flags + NWB_HX_CRIT + wslot, lane, branches);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) branches += __shfl_xor_sync(NWB_FULL_MASK, branches, o);
        if (lane == 0 && branches) atomicAdd(&p.summary->branch_count, branches);
    }
}

#ifndef NWB_EMU
template <bool PUBLISH>
static int nwb_hy_launch_t(const NwbStripParams &sp, const NwbPkConsts &pc, int grid, cudaStream_t st, nwb_fail_fn fail)
{
    auto kernel = nwb_fill_hy_kernel<PUBLISH>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NWB_HX_SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    void *args[] = {(void *)&sp, (void *)&pc};
    e = cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(32 * NWB_HX_WARPS), args, NWB_HX_SMEM_BYTES, st);
    if (e != cudaSuccess) return fail(e, "cudaLaunchCooperativeKernel");
    return 0;
}
static inline int nwb_hy_launch(const NwbStripParams &sp, const NwbPkConsts &pc, int grid, cudaStream_t st, nwb_fail_fn fail)
{
    return sp.publish_rows ? nwb_hy_launch_t<true>(sp, pc, grid, st, fail) : nwb_hy_launch_t<false>(sp, pc, grid, st, fail);
}
#endif
