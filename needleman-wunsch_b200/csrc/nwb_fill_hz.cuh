/*
 * nwb_fill_hz.cuh -- packed 16x2 difference fill with packing AND flush warps ("hz").
 *
 * Same recurrence, strip pipeline, boundary streams and arrow codes as nwb_fill_hx.cuh
 * (the reference's score_cell(), needleman-wunsch.c:418-510), for the case that every
 * strip of the launch can have an SM HALF to itself: at most two strips per SM (tables
 * of up to 2 x 148 x 256 = 75,776 columns per GPU: BASELINE configs 2 and 5 on one GPU,
 * config 3 from two GPUs on).
 *
 * Why.  What bounds a single pair is the strip-to-strip critical path, B/2 + 72 steps
 * per strip, each executed by ONE sweeping warp at ~2 cycles per instruction.
 * In nwb_fill_hx.cuh a step is 78 instructions + ~9 of sub-block overhead, 27 of them the
 * SWAR packing of the row's z, a, u into nibble words.  Here the sweeping warp only runs the
 * recurrence and leaves the raw registers (z, a, u: 24 words per step and lane) in a
 * shared-memory ring; a PACKING warp on another SM sub-partition turns them into the {P1, P2}
 * words of nwb_fill_hx.cuh and feeds the same de-skew ring, which the FLUSH warp (unchanged)
 * turns into arrow codes, branch counts and 128-byte row stores.
 *
 * Block = 8 warps on the four sub-partitions (warp w runs on sub-partition w % 4):
 *     warp 0, 1     sweep strips                         (sub-partitions 0, 1)
 *     warp 2, 3     pack for warp 0, 1                   (sub-partitions 2, 3)
 *     warp 6, 7     flush for warp 0, 1                  (sub-partitions 2, 3)
 *     warp 4, 5     exit at once
 * so a sweeping warp has its sub-partition to itself and its two helpers share another one.
 * Shared memory per sweeping warp: raw ring 16 steps x 3 KB + de-skew ring 128 steps x 512 B.
 *
 * MEASURED SLOWER than nwb_fill_hx.cuh and therefore only in the experiments build (-DNWB_EXPERIMENTS, nwb_tune
 * "pk_hz"): 0.906 vs 0.733 ms at 10k x 10k, 2.67 vs 2.15 ms at 30k x 30k (bit-exact, goldens and whole-table digests).
 * ptxas needs ~20 IMAD.MOV per step to line the raw values up in the aligned register quads of the six STS.128, so
 * the sweeping warp's step does not get shorter (SASS: 78 instructions again), and the raw ring moves 7 KB of shared
 * memory per step and strip.  Kept for the record.
 */
#pragma once
#include "nwb_fill_hx.cuh"

#define NWB_HZ_CRIT 2
#define NWB_HZ_WARPS 8
#define NWB_HZ_RAW_STEPS 16                   /* two sub-blocks of 8 steps */
#define NWB_HZ_RAW_SLOT_BYTES (6 * 512)       /* per step: {z, a, u} x 2 sub-rows, 512 B (32 lanes x 16 B) each */
#define NWB_HZ_RAW_BYTES (NWB_HZ_RAW_STEPS * NWB_HZ_RAW_SLOT_BYTES)
#define NWB_HZ_SLOT_SMEM (NWB_HZ_RAW_BYTES + NWB_HX_RING_BYTES)
#define NWB_HZ_SMEM_BYTES ((size_t)NWB_HZ_CRIT * NWB_HZ_SLOT_SMEM + 64)

/* strips per launch that nwb_fill_hz_kernel can hold (one sweeping warp each, all resident) */
static inline bool nwb_hz_usable(int n_local_strips, int sm_count) { return n_local_strips <= NWB_HZ_CRIT * sm_count; }

__device__ __forceinline__ uint4 nwb_lds128(nwb_smem_addr a)
{
#ifdef NWB_EMU
    return *reinterpret_cast<const uint4 *>(a);
#else
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
#endif
}

/* One row step of one lane of a sweeping warp: the recurrence of nwb_hx_step without the packing.  raw = this
 * lane's 16 bytes of the step's first 512-byte raw plane (planes: z, a, u of sub-row 0, then of sub-row 1). */
template <bool LEAN>
__device__ __forceinline__ void nwb_hz_step(NwbHxState &st, const NwbPkConsts &pc, const NwbPkRange<2> &rg,
                                             const unsigned bq, const int t, const int lane, const int g_idx,
                                             const int g_hi, const int A, const int col_lo, const int col_hi,
                                             const unsigned chars, const nwb_smem_addr raw, uint32_t *out_w,
                                             const bool pub31, unsigned &rs32)
{
    unsigned recv = __shfl_up_sync(NWB_FULL_MASK, st.send, 1);
    const unsigned b = __shfl_sync(NWB_FULL_MASK, bq, t);
    if (lane == 0) recv = b;
    unsigned vL[2];
    vL[0] = __byte_perm(recv, st.vlast[0], 0x5410);
    vL[1] = __byte_perm(recv, st.vlast[1], 0x5432);
    st.sp[0] = __byte_perm(chars, st.sp[0], 0x5410);
    st.sp[1] = __byte_perm(chars, st.sp[1], 0x5432);
    unsigned uafter[2][4];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        unsigned v = vL[r];
        unsigned z[4], a[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned nx = st.tpw[k] ^ st.sp[r];
            a[k] = __viaddmax_s16x2(nx, pc.TT1, pc.AMIS);
            z[k] = __vimax3_s16x2(a[k], v, st.u[k]);
            const unsigned un = z[k] - v;
            const unsigned vn = z[k] - st.u[k];
            st.u[k] = un;
            uafter[r][k] = un;
            v = vn;
        }
        st.vlast[r] = v;
        nwb_sts128(raw + (unsigned)((3 * r + 0) * 512), z[0], z[1], z[2], z[3]);
        nwb_sts128(raw + (unsigned)((3 * r + 1) * 512), a[0], a[1], a[2], a[3]);
        nwb_sts128(raw + (unsigned)((3 * r + 2) * 512), st.u[0], st.u[1], st.u[2], st.u[3]);
    }
    st.send = __byte_perm(st.vlast[0], st.vlast[1], 0x7632);
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
    const bool pub = pub31 && (LEAN || (unsigned)g_hi < rg.ngroups);
    nwb_st_relaxed_sys_pred(out_w + g_idx, st.send | 0x80008000u, pub);
}

/* A sweeping warp's strip.  rsub counts this warp's sub-blocks of 8 steps over all its strips: sub-block n uses
 * half (n & 1) of the raw ring; it may be written once the packing warp is done with sub-block n - 2. */
__device__ __forceinline__ bool nwb_hz_strip(const NwbStripParams &p, const NwbPkConsts &pc, const int c,
                                              unsigned char *raw_ring, volatile int *raw_ready, volatile int *raw_done,
                                              int &rsub, const int lane, long long &rsum)
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

    const nwb_smem_addr lane_raw = nwb_smem_address(raw_ring) + (unsigned)(lane * 16);
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
        const bool lean = !(is_last && R * (s0 + 32) >= B);
#pragma unroll 1
        for (int sub = 0; sub < 32 / NWB_PK_SUB; sub++) {
            const int ss = s0 + NWB_PK_SUB * sub;
            /* raw-ring back-pressure: the packing warp only ever waits for THIS warp, so it always gets there --
             * unless it left because some watchdog fired: look at the error word now and then */
            if (rsub >= 2 && nwb_flag_load(raw_done) < rsub - 1) {
                for (;;) {
                    bool freed = false;
#pragma unroll 1
                    for (int it = 0; it < NWB_WD_POLLS; it++) {
                        if (nwb_flag_load(raw_done) >= rsub - 1) {
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
            if (has_left) {
                const int gs = ss + lane;
                unsigned w = bq_next;
                bool ok = (lane >= NWB_PK_SUB) || (gs >= ngroups) || (w & NWB_PK_VALID) || NWB_DBG_BITS(p, 1);
                if (!__all_sync(NWB_FULL_MASK, ok)) {
                    NwbWatchdog wd; /* bookkeeping outside the inner poll loop: see nwb_fill_hx.cuh */
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
            /* copied out of their load registers before the next loads are issued (counting scoreboard: nwb_fill_hx.cuh) */
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
            const nwb_smem_addr slot0 = lane_raw + (unsigned)((rsub & 1) * (NWB_PK_SUB * NWB_HZ_RAW_SLOT_BYTES));
            if (lean) {
#pragma unroll
                for (int t = 0; t < NWB_PK_SUB; t++)
                    nwb_hz_step<true>(st, pc, rg, bq, t, lane, t, gb + t, A, col_lo, col_hi, chars[t],
                                      slot0 + (unsigned)(t * NWB_HZ_RAW_SLOT_BYTES), outb, pub31, rs32);
            } else {
#pragma unroll
                for (int t = 0; t < NWB_PK_SUB; t++)
                    nwb_hz_step<false>(st, pc, rg, bq, t, lane, t, gb + t, A, col_lo, col_hi, chars[t],
                                       slot0 + (unsigned)(t * NWB_HZ_RAW_SLOT_BYTES), outb, pub31, rs32);
            }
            __syncwarp();
            if (lane == 0) nwb_flag_store(raw_ready, rsub + 1);
            rsub++;
        }
    }
    rsum += (long long)(rs32 & 0xFFFFu) + (long long)(rs32 >> 16);
    return true;
}

/* The packing warp of sweeping warp `wslot`: lane h packs what sweeping lane h left in the raw ring (the row's z, a, u
 * as one nibble per column, the DIAG and UP zero tests: exactly the packing of nwb_hx_step) into the de-skew ring
 * that the flush warp reads.  It walks the same strips, blocks and sub-blocks as its sweeping warp. */
__device__ __forceinline__ void nwb_hz_pack(const NwbStripParams &p, const int wslot, unsigned char *raw_ring,
                                             unsigned char *ring, volatile int *raw_ready, volatile int *raw_done,
                                             volatile int *ready, volatile int *done, const int lane)
{
    const int B = p.B;
    const int nworkers = (int)gridDim.x * NWB_HZ_CRIT;
    const int worker = wslot * (int)gridDim.x + (int)blockIdx.x;
    const int ngroups = (B + 1) / 2;
    const int nblocks = (ngroups + 63 + 31) / 32;
    const nwb_smem_addr lane_raw = nwb_smem_address(raw_ring) + (unsigned)(lane * 16);
    const nwb_smem_addr lane_ring = nwb_smem_address(ring) + (unsigned)(lane * 16);
    int seq = 0, rsub = 0;
    for (int c = p.strip_begin + worker; c < p.strip_end; c += nworkers) {
        unsigned nu = 0u;
        for (int blk = 0; blk < nblocks; blk++) {
            /* de-skew ring back-pressure, as the sweeping warp of nwb_fill_hx.cuh: block seq overwrites the slots of
             * block seq - 4; the rows the flush warp takes after block seq - 2 still read them */
            if (seq >= 2) {
                NwbWatchdog wd;
                while (nwb_flag_load(done) < seq - 1) {
                    nwb_spin_pause(false);
                    if (wd.tick(NWB_ERR_WORD(p), p.watchdog_ns)) return;
                }
            }
#pragma unroll 1
            for (int sub = 0; sub < 32 / NWB_PK_SUB; sub++) {
                {
                    NwbWatchdog wd;
                    while (nwb_flag_load(raw_ready) < rsub + 1) {
                        nwb_spin_pause(false);
                        if (wd.tick(NWB_ERR_WORD(p), p.watchdog_ns)) return; /* the sweeping warp gave up (or never came) */
                    }
                }
#ifndef NWB_EMU
                __threadfence_block();
#endif
                const nwb_smem_addr rslot0 = lane_raw + (unsigned)((rsub & 1) * (NWB_PK_SUB * NWB_HZ_RAW_SLOT_BYTES));
                const nwb_smem_addr dslot0 =
                    lane_ring + (unsigned)(((32 * seq + NWB_PK_SUB * sub) & (NWB_HX_RING_STEPS - 1)) * NWB_HX_SLOT_BYTES);
#pragma unroll
                for (int t = 0; t < NWB_PK_SUB; t++) {
                    unsigned p1[2], p2[2];
#pragma unroll
                    for (int r = 0; r < 2; r++) {
                        const nwb_smem_addr q = rslot0 + (unsigned)(t * NWB_HZ_RAW_SLOT_BYTES + 3 * r * 512);
                        const uint4 z = nwb_lds128(q), a = nwb_lds128(q + 512u), u = nwb_lds128(q + 1024u);
                        const unsigned Z4 = ((z.w * 16u + z.z) * 16u + z.y) * 16u + z.x;
                        const unsigned A4 = ((a.w * 16u + a.z) * 16u + a.y) * 16u + a.x;
                        const unsigned NU = ((u.w * 16u + u.z) * 16u + u.y) * 16u + u.x;
                        const unsigned ZT = A4 - Z4 + NWB_HX_B8; /* bit 3 of a nibble: z == a (DIAG)            */
                        const unsigned ZV = nu - Z4 + NWB_HX_B8; /* vn = z - u(row above); bit 3: vn == 0 (UP) */
                        nu = NU;
                        p1[r] = (ZT & NWB_HX_B8) | NU;
                        p2[r] = ZV;
                    }
                    nwb_sts128(dslot0 + (unsigned)(t * NWB_HX_SLOT_BYTES), p1[0], p2[0], p1[1], p2[1]);
                }
                __syncwarp();
                if (lane == 0) nwb_flag_store(raw_done, rsub + 1);
                rsub++;
            }
            __syncwarp();
            if (lane == 0) nwb_flag_store(ready, seq + 1);
            seq++;
        }
    }
}

template <bool PUBLISH>
__global__ void __launch_bounds__(32 * NWB_HZ_WARPS, 1) nwb_fill_hz_kernel(const NwbStripParams p, const NwbPkConsts pc)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *smem = NWB_SMEM_BASE();
    /* flags per sweeping warp: {raw_ready, raw_done, ready, done} */
    volatile int *flags = reinterpret_cast<volatile int *>(smem + (size_t)NWB_HZ_CRIT * NWB_HZ_SLOT_SMEM);
    if (threadIdx.x < 4 * NWB_HZ_CRIT) flags[threadIdx.x] = 0;
    __syncthreads();
    const int role = warp >> 1; /* 0: sweep (warps 0, 1), 1: pack (2, 3), 2: exit (4, 5), 3: flush (6, 7) */
    const int wslot = warp & 1;
    unsigned char *raw_ring = smem + (size_t)wslot * NWB_HZ_SLOT_SMEM;
    unsigned char *ring = raw_ring + NWB_HZ_RAW_BYTES;
    volatile int *f = flags + 4 * wslot;
    if (role == 0) {
        const int nworkers = (int)gridDim.x * NWB_HZ_CRIT;
        const int worker = wslot * (int)gridDim.x + (int)blockIdx.x;
        long long rsum = 0;
        int rsub = 0;
        for (int c = p.strip_begin + worker; c < p.strip_end; c += nworkers)
            if (!nwb_hz_strip(p, pc, c, raw_ring, f + 0, f + 1, rsub, lane, rsum)) return; /* watchdog */
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(NWB_FULL_MASK, rsum, o);
        if (lane == 0 && rsum) atomicAdd((unsigned long long *)&p.summary->rsum, (unsigned long long)rsum);
    } else if (role == 1) {
        nwb_hz_pack(p, wslot, raw_ring, ring, f + 0, f + 1, f + 2, f + 3, lane);
    } else if (role == 3) {
        unsigned branches = 0;
        nwb_hx_flush<PUBLISH>(p, wslot * (int)gridDim.x + (int)blockIdx.x, (int)gridDim.x * NWB_HZ_CRIT, ring, f + 2, f + 3, lane, branches);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) branches += __shfl_xor_sync(NWB_FULL_MASK, branches, o);
        if (lane == 0 && branches) atomicAdd(&p.summary->branch_count, branches);
    }
}

#ifndef NWB_EMU
template <bool PUBLISH>
static int nwb_hz_launch_t(const NwbStripParams &sp, const NwbPkConsts &pc, int grid, cudaStream_t st, nwb_fail_fn fail)
{
    auto kernel = nwb_fill_hz_kernel<PUBLISH>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NWB_HZ_SMEM_BYTES);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    void *args[] = {(void *)&sp, (void *)&pc};
    e = cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(32 * NWB_HZ_WARPS), args, NWB_HZ_SMEM_BYTES, st);
    if (e != cudaSuccess) return fail(e, "cudaLaunchCooperativeKernel");
    return 0;
}
static inline int nwb_hz_launch(const NwbStripParams &sp, const NwbPkConsts &pc, int grid, cudaStream_t st, nwb_fail_fn fail)
{
    return sp.publish_rows ? nwb_hz_launch_t<true>(sp, pc, grid, st, fail) : nwb_hz_launch_t<false>(sp, pc, grid, st, fail);
}
#endif
