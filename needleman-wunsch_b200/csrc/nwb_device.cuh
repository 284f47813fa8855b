/*
 * nwb_device.cuh -- device-side primitives and launch parameter blocks shared
 * by the fill kernels.  Compiles for sm_100a with nvcc and, with -DNWB_EMU,
 * for the host under the test-only SIMT emulator (tests/emu/emu_cuda.h).
 */
#pragma once

#include <stdint.h>
#include <stddef.h>

#ifdef NWB_EMU
#include "emu_cuda.h"
#define NWB_SMEM_BASE() (emu_smem())
#else
#include <cuda_runtime.h>
extern __shared__ __align__(16) unsigned char nwb_dyn_smem_[];
#define NWB_SMEM_BASE() (nwb_dyn_smem_)
#endif

#define NWB_FULL_MASK 0xffffffffu

/* ---- inter-block / inter-GPU flag protocol --------------------------------
 * Producer: plain stores of the boundary data by ONE thread, then
 * st.release of the progress word by the same thread.  Consumer: every lane
 * ld.acquire's the progress word until it is large enough, then reads the
 * data.  `_sys` variants are used when producer and consumer sit on
 * different GPUs (peer memory over NVLink). */
__device__ __forceinline__ int nwb_ld_acquire_gpu(const int *p)
{
#ifdef NWB_EMU
    return *(const volatile int *)p;
#else
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ int nwb_ld_acquire_sys(const int *p)
{
#ifdef NWB_EMU
    return *(const volatile int *)p;
#else
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void nwb_st_release_gpu(int *p, int v)
{
#ifdef NWB_EMU
    *(volatile int *)p = v;
#else
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ void nwb_st_release_sys(int *p, int v)
{
#ifdef NWB_EMU
    *(volatile int *)p = v;
#else
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ void nwb_pause()
{
    __nanosleep(32);
}

/* shared-memory stores through an explicit shared-window address (keeps the
 * address arithmetic 32-bit and out of the generic address space) */
#ifdef NWB_EMU
typedef unsigned char *nwb_smem_addr;
__device__ __forceinline__ nwb_smem_addr nwb_smem_address(unsigned char *p) { return p; }
template <typename T>
__device__ __forceinline__ void nwb_sts(nwb_smem_addr a, int off, T v) { *reinterpret_cast<T *>(a + off) = v; }
#else
typedef unsigned nwb_smem_addr;
__device__ __forceinline__ nwb_smem_addr nwb_smem_address(unsigned char *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <typename T>
__device__ __forceinline__ void nwb_sts(nwb_smem_addr a, int off, T v);
template <>
__device__ __forceinline__ void nwb_sts<uint32_t>(nwb_smem_addr a, int off, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a + (unsigned)off), "r"(v));
}
template <>
__device__ __forceinline__ void nwb_sts<uint16_t>(nwb_smem_addr a, int off, uint16_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a + (unsigned)off), "h"(v));
}
template <>
__device__ __forceinline__ void nwb_sts<uint8_t>(nwb_smem_addr a, int off, uint8_t v)
{
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a + (unsigned)off), "r"((unsigned)v));
}
#endif

__device__ __forceinline__ unsigned long long nwb_globaltimer()
{
#ifdef NWB_EMU
    return 0ull;
#else
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
#endif
}

/* A register copy the compiler can neither elide nor move across memory operations (see nwb_fill_hx.cuh). */
__device__ __forceinline__ unsigned nwb_pin_copy(unsigned v)
{
#ifdef NWB_EMU
    return v;
#else
    unsigned r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v) : "memory");
    return r;
#endif
}

/* read-only (non-coherent) 16-bit load */
__device__ __forceinline__ unsigned short nwb_ldg_u16(const uint16_t *p)
{
#ifdef NWB_EMU
    return *p;
#else
    return __ldg(p);
#endif
}

__device__ __forceinline__ unsigned nwb_ldg_u32(const unsigned *p)
{
#ifdef NWB_EMU
    return *p;
#else
    return __ldg(p);
#endif
}

/* relaxed (L2-coherent, L1-bypassing) 32-bit accesses for self-validating stream words */
__device__ __forceinline__ unsigned nwb_ld_relaxed_u32(const uint32_t *p, bool sys)
{
#ifdef NWB_EMU
    (void)sys;
    return *(const volatile uint32_t *)p;
#else
    unsigned v;
    if (sys) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void nwb_st_relaxed_u32(uint32_t *p, unsigned v, bool sys)
{
#ifdef NWB_EMU
    (void)sys;
    *(volatile uint32_t *)p = v;
#else
    if (sys) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}

/* Spin until *flag >= need, then order the following data reads after it.  The
 * poll itself is a relaxed (L2) load: an ld.acquire inside the loop costs an L1
 * invalidation (CCTL.IVALL) per iteration.  `sys` selects system scope (peer memory). */
__device__ __forceinline__ void nwb_wait_ge(const int *flag, int need, bool sys)
{
#ifdef NWB_EMU
    while (*(const volatile int *)flag < need) nwb_pause();
#else
    int v;
    if (sys) {
        do {
            asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag));
        } while (v < need);
        asm volatile("fence.acq_rel.sys;" ::: "memory");
    } else {
        do {
            asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag));
        } while (v < need);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
#endif
}

/* ---- watchdog for device-side spin loops -------------------------------------
 * Every wait on another warp / block / GPU is bounded: tick() is called once per poll iteration from a
 * converged warp; every 1024 polls it looks at the launch's error word and at %globaltimer.  It returns true
 * (for the whole warp) when the wait must be abandoned: nothing arrived within limit_ns -- then the error word is
 * set, so that every other waiting warp leaves too and the host reports NWB_ERR_CUDA instead of hanging the GPU
 * (a lost boundary word, a partner kernel that is not co-resident) -- or some other warp already gave up. */
#define NWB_DEVERR_WATCHDOG 1
#ifdef NWB_EXPERIMENTS
#define NWB_DBG_BITS(p, bits) ((p).debug_nowait & (bits)) /* experiments build: skip waits (results are wrong) */
#else
#define NWB_DBG_BITS(p, bits) 0
#endif
#define NWB_FAULT_INJECTED(p) (((p).debug_nowait & 4) != 0) /* test only: boundary streams are not published */

#define NWB_ERR_WORD(p) ((p).summary ? &(p).summary->error : (int *)0)
#ifndef NWB_WD_MODE
#define NWB_WD_MODE 2
#endif
#define NWB_WD_POLLS 1024 /* polls between two looks at the clock */
struct NwbWatchdog {
    unsigned polls;
    unsigned long long t0;
    __device__ __forceinline__ NwbWatchdog() : polls(0u), t0(0ull) {}
#if NWB_WD_MODE == 1
    __device__ __noinline__ bool slow(int *err, unsigned long long limit_ns)
#else
    __device__ __forceinline__ bool slow(int *err, unsigned long long limit_ns)
#endif
    {
        if (!err) return false; /* a launch without a summary block (batch kernels: no cross-warp waits) */
        bool give_up = (*reinterpret_cast<volatile int *>(err) != 0);
        const unsigned long long now = nwb_globaltimer();
        if (t0 == 0ull) t0 = now | 1ull;
        else if (now - t0 > limit_ns) {
            atomicExch(err, NWB_DEVERR_WATCHDOG);
            give_up = true;
        }
        return __any_sync(NWB_FULL_MASK, give_up) != 0;
    }
    __device__ __forceinline__ bool tick(int *err, unsigned long long limit_ns)
    {
#if NWB_WD_MODE == 0
        (void)err; (void)limit_ns;
        return false;
#else
        if ((++polls & 1023u) != 0u) return false;
        return slow(err, limit_ns);
#endif
    }
};

/* nwb_wait_ge with the watchdog: false = the wait was abandoned (all lanes of the warp call this together) */
__device__ __forceinline__ bool nwb_wait_ge_wd(const int *flag, int need, bool sys, int *err, unsigned long long limit_ns)
{
    NwbWatchdog wd;
    for (;;) {
        bool arrived = false;
#pragma unroll 1
        for (int it = 0; it < NWB_WD_POLLS; it++) {
#ifdef NWB_EMU
            const int v = *(const volatile int *)flag;
#else
            int v;
            if (sys) asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag));
            else asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag));
#endif
            if (__all_sync(NWB_FULL_MASK, v >= need)) {
                arrived = true;
                break;
            }
#ifdef NWB_EMU
            nwb_pause();
#endif
        }
        if (arrived) break;
        if (wd.slow(err, limit_ns)) return false;
    }
#ifndef NWB_EMU
    if (sys) asm volatile("fence.acq_rel.sys;" ::: "memory");
    else asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
    return true;
}

/* ---- 16x2 helpers (DPX / video SIMD; VIMNMX.U16x2, VIMNMX3.U16x2, VIADD.16x2) */
__device__ __forceinline__ unsigned nwb_min_u16x2(unsigned a, unsigned b) { return __vminu2(a, b); }
__device__ __forceinline__ unsigned nwb_min3_u16x2(unsigned a, unsigned b, unsigned c) { return __vimin3_u16x2(a, b, c); }
__device__ __forceinline__ unsigned nwb_max3_u16x2(unsigned a, unsigned b, unsigned c) { return __vimax3_u16x2(a, b, c); }

/* ---- results ---------------------------------------------------------------- */
struct NwbDevSummary {
    int opt_score;
    unsigned branch_count;
    int greatest_abs;
    int kernel_kind;
    unsigned long long count;
    long long rsum; /* packed kernel: sum of bottom-row u differences */
    int count_state;      /* NWB_SPC_* (nwb_count_sparse.cuh): 1 = the sparse backward sweep produced `count` */
    unsigned sparse_rows; /* rows the sparse sweep visited before its live set died                            */
    int error;            /* != 0: a device-side watchdog gave up waiting (NWB_DEVERR_*); results are invalid  */
    int ticket;           /* nwb_fill_hx_kernel: the next block id (blocks draw their ids as they start)                */
    unsigned long long dig_row; /* dense count sweep with digests: sum_i mix64(i, cnt(i,B)) over this rank's columns */
    unsigned long long dig_col; /* ... sum_j mix64(j, cnt(A,j)) (the rank that owns column A)                        */
};

/* ---- launch parameters of the single-pair strip-pipeline kernels ------------
 * The table is cut into column strips of `strip_w` interior columns; strip c
 * covers columns i in [c*strip_w + 1, (c+1)*strip_w].  One warp sweeps one
 * strip top to bottom.  Strip c streams its right boundary column into
 * bnd_*[c - strip_begin] and publishes in progress[c - strip_begin] how many
 * rows are final; strip c+1 polls it.  Strips [strip_begin, strip_end) belong
 * to this launch (this GPU).  When the table starts on another GPU
 * (strip_begin > 0) the boundary of strip strip_begin-1 arrives in the in_*
 * "inbox", written by the left-neighbour GPU through peer memory; when the
 * table continues on another GPU (strip_end < n_strips) the last local strip
 * writes into out_* = the right neighbour's inbox. */
struct NwbStripParams {
    const uint8_t *top;  /* A bytes  */
    const uint8_t *side; /* B bytes  */
    const uint16_t *side_pre; /* packed kernel: pre-shifted, complemented, padded side string */
    int A, B;
    int m, k, d;
    int n_strips;     /* total strips of the table                         */
    int strip_begin;  /* first strip of this launch                        */
    int strip_end;    /* one past the last strip of this launch            */
    uint8_t *arrows;  /* nibble table, B rows x pitch bytes                */
    size_t pitch;
    int32_t *scores;        /* interior scores, B rows x spitch (or NULL)  */
    unsigned long long *cntmat; /* interior counts, B rows x spitch (NULL) */
    size_t spitch;          /* elements per row of scores / cntmat         */
    /* boundary streams of the local strips, [c - strip_begin][row j] */
    int32_t *bnd_s;             /* general kernel: int32 score of (c_last, j) */
    unsigned long long *bnd_c;  /* count of (c_last, j) (WANT_COUNT)          */
    uint32_t *bnd_w;            /* packed kernel: stream word per row         */
    size_t bpitch;              /* elements per strip in bnd_*                */
    int *progress;              /* [strip_end - strip_begin] rows published   */
    /* inbox (valid iff strip_begin > 0): written by the left-neighbour GPU */
    const int32_t *in_bnd_s;
    const unsigned long long *in_bnd_c;
    const uint32_t *in_bnd_w;
    const int *in_progress;
    /* right neighbour's inbox (valid iff strip_end < n_strips) */
    int32_t *out_bnd_s;
    unsigned long long *out_bnd_c;
    uint32_t *out_bnd_w;
    int *out_progress;
    NwbDevSummary *summary;
    int count_branches; /* packed kernel: count cells with >= 2 arrows while flushing rows */
    const uint32_t *gate_ack; /* hx kernel in queue mode, pipelined strip group: the right neighbour's acknowledgement word ... */
    unsigned gate_need;       /* ... and the value it must have reached before this launch writes into the neighbour's inbox   */
    int hx_spb;         /* hx kernel: 0 = strips dealt out cyclically over a resident grid (one fill as fast as possible);
                         * 1..3 = queue mode, blocks draw tickets and sweep that many adjacent strips (nwb_fill_hx.cuh) */
    int publish_rows;   /* hx kernel: publish in progress[] how many arrow rows of each strip are in memory
                         * (the count sweep of nwb_count.cuh trails the fill on a second stream) */
    int debug_nowait; /* bit 2 (value 4): fault injection for the watchdog test -- the strips' boundary streams are
                       * not published, so the next strip's wait times out (NWB_ERR_CUDA).  Bits 0, 1 (experiments
                       * build only): skip the inter-strip waits / the flush (results are wrong) */
    unsigned long long watchdog_ns; /* a spin loop that sees no progress for this long sets summary->error and leaves */
    unsigned long long *debug_times; /* diagnostics: per strip {entry, first words valid, step 64, exit} in ns, or NULL */
    unsigned long long *debug_trace; /* diagnostics: [8 traced strips][nblocks][2] = {ns at block start, polls so far} */
    int debug_trace_stride;          /* strips c with c % stride == 0 are traced (slot c / stride, < 8) */
    int debug_trace_blocks;
};
