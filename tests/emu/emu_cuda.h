/*
 * emu_cuda.h -- a tiny SIMT emulator so that the CUDA kernel sources under
 * needleman-wunsch_b200/csrc/ can be executed on the GPU-less development box.
 *
 * TEST INFRASTRUCTURE ONLY.  It is never part of libnwb.so: the product path
 * has no CPU fallback.  tests/test_emu_*.py compile the kernel headers with
 * -DNWB_EMU against this file and compare against the oracle, which catches
 * indexing / protocol bugs before GPU minutes are spent.  It does NOT model
 * the memory model (fences are no-ops, every access is sequentially
 * consistent), so races must still be checked on the GPU (compute-sanitizer).
 *
 * Model: every CUDA thread is a fiber (own stack, hand-rolled x86-64 context
 * switch) run by one OS thread, round-robin.  A fiber runs until it reaches a
 * warp/block collective (shuffle, __syncwarp, __syncthreads) or a spin-wait
 * pause, where it yields.
 */
#ifndef EMU_CUDA_H
#define EMU_CUDA_H

#if !defined(__x86_64__)
#error "the SIMT emulator's context switch is x86-64 only"
#endif

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <functional>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __restrict__
#define __noinline__ __attribute__((noinline))

struct emu_dim3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 v; v.x = x; v.y = y; return v; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

struct emu_warp {
    uint64_t slot[32];
    int arrived;
    unsigned gen;
    int nlanes;
};

struct emu_block {
    unsigned char *smem;
    emu_warp *warps;
    int nwarps;
    int bar_arrived;
    unsigned bar_gen;
    int nthreads;
    int live;
};

struct emu_thread {
    void *sp;
    void *stack;
    emu_dim3 tid, bid;
    emu_block *block;
    emu_warp *warp;
    int lane;
    int done;
};

extern emu_thread *emu_cur;
extern emu_dim3 emu_blockDim, emu_gridDim;
void emu_yield(void);
/* Launch: runs body() once per thread of a grid x block launch. */
void emu_launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &body);

#define threadIdx (emu_cur->tid)
#define blockIdx (emu_cur->bid)
#define blockDim emu_blockDim
#define gridDim emu_gridDim

static inline unsigned char *emu_smem(void) { return emu_cur->block->smem; }

/* ---- warp collectives ---------------------------------------------------- */
static inline void emu_warp_barrier(void)
{
    emu_warp *w = emu_cur->warp;
    const unsigned g = w->gen;
    if (++w->arrived == w->nlanes) {
        w->arrived = 0;
        w->gen = g + 1;
    } else {
        while (w->gen == g) emu_yield();
    }
}

static inline uint64_t emu_exchange(uint64_t v, int src_lane)
{
    emu_warp *w = emu_cur->warp;
    w->slot[emu_cur->lane] = v;
    emu_warp_barrier();
    const uint64_t r = w->slot[src_lane & 31];
    emu_warp_barrier();
    return r;
}

template <typename T>
static inline T __shfl_sync(unsigned, T v, int src, int = 32)
{
    uint64_t x = 0;
    memcpy(&x, &v, sizeof(T));
    x = emu_exchange(x, src);
    T r;
    memcpy(&r, &x, sizeof(T));
    return r;
}
template <typename T>
static inline T __shfl_up_sync(unsigned m, T v, unsigned delta, int = 32)
{
    const int lane = emu_cur->lane;
    const int src = lane - (int)delta;
    T r = __shfl_sync(m, v, src < 0 ? lane : src);
    return r;
}
template <typename T>
static inline T __shfl_down_sync(unsigned m, T v, unsigned delta, int = 32)
{
    const int lane = emu_cur->lane;
    const int src = lane + (int)delta;
    return __shfl_sync(m, v, src > 31 ? lane : src);
}
template <typename T>
static inline T __shfl_xor_sync(unsigned m, T v, int mask, int = 32)
{
    return __shfl_sync(m, v, emu_cur->lane ^ mask);
}
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp_barrier(); }
static inline int __all_sync(unsigned, int pred)
{
    emu_warp *w = emu_cur->warp;
    w->slot[emu_cur->lane] = pred ? 1 : 0;
    emu_warp_barrier();
    int all = 1;
    for (int i = 0; i < w->nlanes; i++) all &= (int)w->slot[i];
    emu_warp_barrier();
    return all;
}
static inline int __any_sync(unsigned m, int pred) { return !__all_sync(m, !pred); }
static inline void __syncthreads(void)
{
    emu_block *b = emu_cur->block;
    const unsigned g = b->bar_gen;
    if (++b->bar_arrived == b->nthreads) {
        b->bar_arrived = 0;
        b->bar_gen = g + 1;
    } else {
        while (b->bar_gen == g) emu_yield();
    }
}
static inline void __threadfence(void) {}
static inline void __threadfence_block(void) {}
static inline void __threadfence_system(void) {}
static inline void __nanosleep(unsigned) { emu_yield(); }

/* ---- integer intrinsics --------------------------------------------------- */
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)((v >> (8 * (sel & 7))) & 0xff);
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh)
{
    sh &= 31;
    return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh)
{
    sh &= 31;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
static inline int __vimax3_s32(int a, int b, int c) { int m = a > b ? a : b; return m > c ? m : c; }
static inline int __vimin3_s32(int a, int b, int c) { int m = a < b ? a : b; return m < c ? m : c; }
static inline int __viaddmax_s32(int a, int b, int c) { int s = (int)((unsigned)a + (unsigned)b); return s > c ? s : c; }
#define EMU_PER_HALF(expr_lo, expr_hi) ((unsigned)((expr_lo) & 0xffffu) | ((unsigned)((expr_hi) & 0xffffu) << 16))
static inline unsigned emu_min16(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned emu_max16(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned __vminu2(unsigned a, unsigned b)
{
    return EMU_PER_HALF(emu_min16(a & 0xffff, b & 0xffff), emu_min16(a >> 16, b >> 16));
}
static inline unsigned __vmaxu2(unsigned a, unsigned b)
{
    return EMU_PER_HALF(emu_max16(a & 0xffff, b & 0xffff), emu_max16(a >> 16, b >> 16));
}
static inline unsigned __vimin3_u16x2(unsigned a, unsigned b, unsigned c) { return __vminu2(__vminu2(a, b), c); }
static inline unsigned __vimax3_u16x2(unsigned a, unsigned b, unsigned c) { return __vmaxu2(__vmaxu2(a, b), c); }
static inline int emu_s16(unsigned v) { return (int)(int16_t)(v & 0xffff); }
static inline unsigned __viaddmax_s16x2(unsigned a, unsigned b, unsigned c)
{
    /* per half: max((int16)(a + b), c), the add wraps in 16 bits */
    const int lo = emu_s16((a & 0xffff) + (b & 0xffff)), hi = emu_s16((a >> 16) + (b >> 16));
    const int clo = emu_s16(c), chi = emu_s16(c >> 16);
    return EMU_PER_HALF((unsigned)(lo > clo ? lo : clo), (unsigned)(hi > chi ? hi : chi));
}
static inline unsigned __vimax3_s16x2(unsigned a, unsigned b, unsigned c)
{
    int lo = emu_s16(a), hi = emu_s16(a >> 16);
    if (emu_s16(b) > lo) lo = emu_s16(b);
    if (emu_s16(b >> 16) > hi) hi = emu_s16(b >> 16);
    if (emu_s16(c) > lo) lo = emu_s16(c);
    if (emu_s16(c >> 16) > hi) hi = emu_s16(c >> 16);
    return EMU_PER_HALF((unsigned)lo, (unsigned)hi);
}
static inline unsigned __vibmin_u16x2(unsigned a, unsigned b, bool *pred_hi, bool *pred_lo)
{
    *pred_hi = (a >> 16) <= (b >> 16);
    *pred_lo = (a & 0xffff) <= (b & 0xffff);
    return __vminu2(a, b);
}
static inline unsigned __vadd2(unsigned a, unsigned b)
{
    return EMU_PER_HALF((a & 0xffff) + (b & 0xffff), (a >> 16) + (b >> 16));
}
static inline unsigned __vsub2(unsigned a, unsigned b)
{
    return EMU_PER_HALF((a & 0xffff) - (b & 0xffff), (a >> 16) - (b >> 16));
}

/* ---- atomics (single OS thread: plain RMW) ------------------------------- */
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }

#endif /* EMU_CUDA_H */
