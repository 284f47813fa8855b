/*
 * nwb_api.cu -- the C ABI of include/nwb.h over the sm_100a fill kernels.
 *
 * Host side only: geometry, device memory, launches, host<->device copies.
 * There is no CPU implementation of the fill in this library; without a CUDA
 * device every entry point fails with NWB_ERR_NO_DEVICE.
 */
#include "../../include/nwb.h"

#include <cuda_runtime.h>

#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "nwb_layout.h"
#include "nwb_fill_i32.cuh"
#include "nwb_fill_pk.cuh"
#include "nwb_fill_hx.cuh"
#ifdef NWB_EXPERIMENTS
/* built, verified and measured SLOWER than nwb_fill_hx.cuh; kept for the record, not in the product build */
#include "nwb_fill_hy.cuh" /* one row of skew per virtual lane */
#include "nwb_fill_hz.cuh" /* packing warps next to the sweeping and flush warps, two strips per SM */
#endif
#include "nwb_count.cuh"
#include "nwb_count_sparse.cuh"
#include "nwb_digest.cuh"
#include "nwb_batch.cuh"
#include "nwb_batch_bx.cuh"
#include "nwb_batch_bp.cuh"
#include "nwb_batch_lcount.cuh"
#include "nwb_batch_count.cuh"
#include "nwb_batch_i32.cuh"
#include "nwb_peak.cuh"

#define NWB_ABI_VERSION 2

/* ------------------------------------------------------------------------- */
static thread_local char g_cuda_err[256] = "";

static int cuda_fail(cudaError_t e, const char *what)
{
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
    if (e == cudaErrorMemoryAllocation) return NWB_ERR_NOMEM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return NWB_ERR_NO_DEVICE;
    return NWB_ERR_CUDA;
}
#define CK(call)                                              \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);   \
    } while (0)

extern "C" const char *nwb_strerror(int err)
{
    switch (err) {
    case NWB_OK: return "ok";
    case NWB_ERR_INVALID: return "invalid argument";
    case NWB_ERR_NOMEM: return "out of memory";
    case NWB_ERR_CUDA: return "CUDA failure";
    case NWB_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
    case NWB_ERR_UNSUPPORTED: return "unsupported flag combination";
    default: return "unknown error";
    }
}
extern "C" const char *nwb_last_cuda_error(void) { return g_cuda_err; }
extern "C" int nwb_abi_version(void) { return NWB_ABI_VERSION; }
extern "C" int nwb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

/* ---- explicit tuning / test overrides (include/nwb.h: nwb_tune) ---------------
 * The library reads NO environment variables.  Every override selects between
 * kernels that produce identical results, except `inject_fault`, which makes a
 * fill fail with NWB_ERR_CUDA through the device watchdog (never a wrong result). */
struct NwbTune {
    int pk_k = 0;        /* 0 = auto; 1, 2, 4 columns per half-lane (nwb_fill_pk.cuh)                       */
    int pk_r = 0;        /* 0 = auto; 1, 2 rows per step                                                     */
    int pk_warps = 0;    /* 0 = auto; sweeping warps per block of nwb_fill_pk_kernel                         */
    int pk_hx = -1;      /* -1 = auto; 0 = never nwb_fill_hx_kernel; 1 = whenever the scheme allows          */
    int count_mode = 0;  /* 0 = auto (sparse backward sweep, dense sweep behind it); 1 = fused into the fill;
                          * 2 = dense sweep after the fill; 3 = dense sweep trailing the fill on a 2nd stream */
    int cnt_cpl = 0;     /* 0 = auto; 2, 4, 8 columns per lane of the dense count sweep                      */
    int batch_bx = -1;   /* -1 = auto; 0 = never nwb_batch_bx_kernel                                         */
    int batch_cx = -1;   /* -1 = auto; 0 = never nwb_batch_cx_kernel                                         */
    int bcnt_chain = -1; /* -1 = auto; 0 = never nwb_batch_count_chain_kernel                                */
    int cx_warps = 0;    /* 0 = auto (12); 16                                                                */
    int batch_bp = -1;   /* -1 = auto; 0 = never nwb_batch_bp_kernel (bit-parallel rows, one thread per pair)      */
    int bp_warps = 0;    /* 0 = auto; warps per block of nwb_batch_bp_kernel (1..16)                          */
    int bp_aligned = -1; /* -1 = auto; 0 = never the 64 KB-aligned look-up table (one PRMT per look-up address)     */
    int batch_lcount = -1; /* -1 = auto; 0 = never nwb_batch_lcount_kernel (sparse count, one thread per pair)      */
    int lc_warps = 0;    /* 0 = auto; warps per block of nwb_batch_lcount_kernel                              */
    int watchdog_ms = 4000; /* device-side spin loops give up after this long without progress               */
    int inject_fault = 0;   /* test only: 1 = the fill's strips do not publish their boundary streams        */
    int plan_cache = 1;     /* 0 = nwb_fill()/nwb_fill_on() create and destroy their device workspace per call */
    int hx_spb = 0;         /* 1..3: the hx kernel in queue mode (see NWB_QUEUE) with that many adjacent strips per block */
#ifdef NWB_EXPERIMENTS
    int pk_hy = 0;
    int pk_hz = 0;
    int debug_nowait = 0;
#endif
};
static NwbTune g_tune;

/* Page-locked host memory for callers that want their inputs to move at full PCIe / C2C speed (a pageable buffer is
 * staged by the driver at ~12 GB/s on this class of host; a pinned one reaches ~50 GB/s). */
extern "C" void *nwb_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (nwb_device_count() <= 0) return nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void nwb_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

extern "C" int nwb_tune(const char *key, int value)
{
    if (!key) return NWB_ERR_INVALID;
    struct { const char *name; int *slot; } tab[] = {
        {"pk_k", &g_tune.pk_k}, {"pk_r", &g_tune.pk_r}, {"pk_warps", &g_tune.pk_warps}, {"pk_hx", &g_tune.pk_hx},
        {"count_mode", &g_tune.count_mode}, {"cnt_cpl", &g_tune.cnt_cpl}, {"batch_bx", &g_tune.batch_bx},
        {"batch_cx", &g_tune.batch_cx}, {"bcnt_chain", &g_tune.bcnt_chain},
        {"cx_warps", &g_tune.cx_warps}, {"batch_bp", &g_tune.batch_bp}, {"bp_warps", &g_tune.bp_warps},
        {"bp_aligned", &g_tune.bp_aligned}, {"batch_lcount", &g_tune.batch_lcount}, {"lc_warps", &g_tune.lc_warps},
        {"watchdog_ms", &g_tune.watchdog_ms}, {"inject_fault", &g_tune.inject_fault}, {"plan_cache", &g_tune.plan_cache},
        {"hx_spb", &g_tune.hx_spb},
#ifdef NWB_EXPERIMENTS
        {"pk_hy", &g_tune.pk_hy}, {"pk_hz", &g_tune.pk_hz}, {"debug_nowait", &g_tune.debug_nowait},
#endif
    };
    for (auto &t : tab)
        if (strcmp(t.name, key) == 0) {
            *t.slot = value;
            return NWB_OK;
        }
    return NWB_ERR_INVALID;
}
extern "C" void nwb_tune_reset(void) { g_tune = NwbTune(); }

/* ------------------------------------------------------------------------- */
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0; /* elements */
    int ensure(size_t n)
    {
        if (n <= cap) return NWB_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc((void **)&p, n * sizeof(T));
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        cap = n;
        return NWB_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

/* inbox: the boundary stream of the strip left of this plan's first strip */
struct Inbox {
    size_t bpitch;      /* elements per array                        */
    size_t off_s, off_c, off_w, off_flag; /* byte offsets in `base`  */
    size_t bytes;
    unsigned char *base;
};

struct NwbIpcBlob {
    cudaIpcMemHandle_t handle;
    unsigned long long bpitch, off_s, off_c, off_w, off_flag, bytes;
    int device;
    int pad;
};

struct nwb_plan {
    int device = 0;
    unsigned flags = 0;
    int maxA = 0, maxB = 0;
    int rank = 0, world = 1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t stream2 = nullptr;           /* the count sweep, when it trails the fill */
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool timed = false;
    int A = 0, B = 0;
    DevBuf<uint8_t> top, side, arrows;
    DevBuf<int32_t> scores, bnd_s;
    DevBuf<unsigned long long> cntmat, bnd_c;
    DevBuf<uint32_t> bnd_w;
    DevBuf<uint16_t> side_pre;
    DevBuf<unsigned long long> digest;
    int count_path = 0;
    bool suppress_count = false; /* nwb_fill_on(): a strip group fills first and tries the sparse count on its last rank */
    DevBuf<int> progress;
    DevBuf<NwbDevSummary> summary;
    Inbox inbox = {};
    /* Pipelined runs (nwb_plan_run_pipelined): the inbox allocation holds TWO copies of the inbox arrays (run e uses
     * copy e & 1) and, behind them, the acknowledgement word: the number of pipelined runs whose inbox copy this rank
     * has finished with and zeroed again. */
    long long epoch = 0;
    /* right neighbour's inbox (peer memory), if attached */
    unsigned char *right_base = nullptr;
    bool right_is_ipc = false;
    Inbox right = {};
    NwbLayout L = {};
    int strip_begin = 0, strip_end = 0;
    int kind = NWB_KIND_I32;
    int64_t launches = 0;
    NwbDevSummary last = {};
    bool ran = false;
    cudaStream_t last_stream = nullptr;
    int m = 0, k = 0, d = 0;
    bool pk_hx = false; /* packed kernel variant with flush warps (nwb_fill_hx.cuh) */
    bool pk_hy = false; /* ... with one row of skew per virtual lane (nwb_fill_hy.cuh) */
    bool pk_hz = false; /* ... with packing warps as well, two strips per SM (nwb_fill_hz.cuh) */
    bool count_pass = false; /* the count runs as a second sweep over the arrow codes (nwb_count.cuh) */
};

/* the allocation behind an inbox: two copies of its arrays (pipelined runs alternate) + the acknowledgement word */
#define NWB_INBOX_ACK_BYTES 256
#define NWB_INBOX_ALLOC_BYTES(ib) (2 * (ib).bytes + NWB_INBOX_ACK_BYTES)
#define NWB_INBOX_ACK(ib_base, ib) (reinterpret_cast<unsigned *>((ib_base) + 2 * (ib).bytes))

static void make_inbox_layout(Inbox &ib, size_t bpitch)
{
    ib.bpitch = bpitch;
    ib.off_s = 0;
    ib.off_c = nwb_round_up(ib.off_s + bpitch * sizeof(int32_t), 256);
    ib.off_w = nwb_round_up(ib.off_c + 2 * bpitch * sizeof(unsigned long long), 256);
    ib.off_flag = nwb_round_up(ib.off_w + bpitch * sizeof(uint32_t), 256);
    ib.bytes = ib.off_flag + 256;
}

extern "C" int nwb_plan_create(int max_top, int max_side, unsigned flags, int device,
                               int strip_rank, int strip_world, nwb_plan **out)
{
    if (!out) return NWB_ERR_INVALID;
    *out = nullptr;
    if (max_top < 0 || max_side < 0 || strip_world < 1 || strip_rank < 0 || strip_rank >= strip_world)
        return NWB_ERR_INVALID;
    if ((flags & (NWB_WANT_SCORES | NWB_WANT_COUNT_MATRIX)) && strip_world > 1) return NWB_ERR_UNSUPPORTED;
    int ndev = nwb_device_count();
    if (ndev <= 0) return NWB_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return NWB_ERR_INVALID;
    CK(cudaSetDevice(device));
    nwb_plan *p = new (std::nothrow) nwb_plan();
    if (!p) return NWB_ERR_NOMEM;
    p->device = device;
    p->flags = flags;
    p->maxA = max_top;
    p->maxB = max_side;
    p->rank = strip_rank;
    p->world = strip_world;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->stream2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&p->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&p->ev1);
    if (e != cudaSuccess) {
        int rc = cuda_fail(e, "plan setup");
        nwb_plan_destroy(p);
        return rc;
    }
    p->sm_count = prop.multiProcessorCount;
    int rc = p->summary.ensure(1);
    if (rc == NWB_OK) rc = p->top.ensure((size_t)max_top + 1024);
    if (rc == NWB_OK) rc = p->side.ensure((size_t)max_side + 1024);
    if (rc == NWB_OK && strip_world > 1) {
        /* inbox sized for the longest side string; allocated once so that it
         * can be exported through CUDA IPC before any fill */
        make_inbox_layout(p->inbox, nwb_round_up((size_t)max_side + 1 + 64 + 512, 32));
        e = cudaMalloc((void **)&p->inbox.base, NWB_INBOX_ALLOC_BYTES(p->inbox));
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMalloc(inbox)");
        else {
            /* on the plan's own (non-blocking) stream, then wait: nothing else orders it against the first run */
            e = cudaMemsetAsync(p->inbox.base, 0, NWB_INBOX_ALLOC_BYTES(p->inbox), p->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
        }
        if (rc == NWB_OK && e != cudaSuccess) rc = cuda_fail(e, "cudaMemset(inbox)");
    }
    if (rc != NWB_OK) {
        nwb_plan_destroy(p);
        return rc;
    }
    *out = p;
    return NWB_OK;
}

extern "C" void nwb_plan_destroy(nwb_plan *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    p->top.release(); p->side.release(); p->arrows.release(); p->scores.release();
    p->bnd_s.release(); p->cntmat.release(); p->bnd_c.release(); p->bnd_w.release();
    p->progress.release(); p->summary.release(); p->side_pre.release(); p->digest.release();
    if (p->inbox.base) cudaFree(p->inbox.base);
    if (p->right_base && p->right_is_ipc) cudaIpcCloseMemHandle(p->right_base);
    if (p->stream2) { cudaStreamSynchronize(p->stream2); cudaStreamDestroy(p->stream2); }
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_join) cudaEventDestroy(p->ev_join);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

extern "C" int nwb_plan_upload(nwb_plan *p, const char *top, int top_len, const char *side, int side_len)
{
    if (!p || top_len < 0 || side_len < 0 || (top_len && !top) || (side_len && !side)) return NWB_ERR_INVALID;
    if (top_len > p->maxA || side_len > p->maxB) return NWB_ERR_INVALID;
    CK(cudaSetDevice(p->device));
    if (top_len) CK(cudaMemcpyAsync(p->top.p, top, (size_t)top_len, cudaMemcpyHostToDevice, p->stream));
    if (side_len) CK(cudaMemcpyAsync(p->side.p, side, (size_t)side_len, cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    p->A = top_len;
    p->B = side_len;
    p->ran = false;
    return NWB_OK;
}

/* ---- kernel selection -------------------------------------------------------
 * The packed 16x2 difference kernel needs every per-cell difference to stay
 * small (nwb_fill_pk.cuh); otherwise, or when scores/abs/count-matrix are
 * requested, the general int32 kernel runs. */
static int choose_kind(unsigned flags, int m, int k, int d, NwbPkConsts *pc)
{
    if (flags & (NWB_FORCE_GENERAL | NWB_WANT_SCORES | NWB_TRACK_ABS | NWB_WANT_COUNT_MATRIX)) return NWB_KIND_I32;
    if (!nwb_pk_supported(m, k, d, pc)) return NWB_KIND_I32;
    return NWB_KIND_PK;
}

template <typename KernelT>
static int launch_strip_kernel(KernelT kernel, int grid, int block, size_t smem, cudaStream_t stream,
                               const NwbStripParams &sp)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
    /* cooperative launch: not for grid.sync, only for its guarantee that all
     * blocks are co-resident (the strips spin-wait on one another) */
    void *args[] = {(void *)&sp};
    e = cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(block), args, smem, stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchCooperativeKernel");
    return NWB_OK;
}

static int run_i32(nwb_plan *p, unsigned flags, const NwbStripParams &sp, int grid, cudaStream_t st)
{
    (void)p;
    const bool C = flags & NWB_WANT_COUNT, S = flags & NWB_WANT_SCORES, AB = flags & NWB_TRACK_ABS,
               CM = flags & NWB_WANT_COUNT_MATRIX;
    const int block = 32 * NWB_I32_WARPS;
    const size_t smem = NWB_I32_SMEM_BYTES;
#define L_(c, s, a, cm) return launch_strip_kernel(nwb_fill_i32_kernel<c, s, a, cm>, grid, block, smem, st, sp)
    if (CM) {
        if (S) { if (AB) L_(true, true, true, true); else L_(true, true, false, true); }
        else { if (AB) L_(true, false, true, true); else L_(true, false, false, true); }
    }
    if (C) {
        if (S) { if (AB) L_(true, true, true, false); else L_(true, true, false, false); }
        else { if (AB) L_(true, false, true, false); else L_(true, false, false, false); }
    }
    if (S) { if (AB) L_(false, true, true, false); else L_(false, true, false, false); }
    if (AB) L_(false, false, true, false);
    L_(false, false, false, false);
#undef L_
}

static int run_pk(nwb_plan *p, const NwbStripParams &sp, const NwbPkConsts &pc, bool count, int grid, int warps,
                  cudaStream_t st);

/* zero `n` 16-byte words unless *state == NWB_SPC_DONE (the dense count sweep's streams, when the sparse
 * backward sweep has not already produced the count) */
__global__ void __launch_bounds__(256) nwb_zero_unless_done_kernel(const int *state, uint4 *dst, size_t n)
{
    if (*reinterpret_cast<const volatile int *>(state) == NWB_SPC_DONE) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = make_uint4(0u, 0u, 0u, 0u);
}

/* where the count behind -s comes from (nwb_summary.count_path) */
enum { NWB_CNT_NONE = 0, NWB_CNT_FUSED = 1, NWB_CNT_DENSE = 2, NWB_CNT_SPARSE = 3, NWB_CNT_SPARSE_BAILED = 4 };

/* ---- pipelined strip groups: flow control between neighbouring ranks ---------------------------------------
 * Consecutive fills of a strip group overlap: rank r starts fill e + 1 while the ranks to its right are still on
 * fill e.  Fill e streams into copy e & 1 of the right neighbour's inbox; the neighbour zeroes that copy when its
 * own fill e is done and then posts e + 1 in its acknowledgement word (stream order: fill, memset, this kernel).
 * Before fill e + 2 the left rank's stream waits in the gate kernel until the word says so -- in steady state it
 * already does.  The wait is watchdog-bounded like every other device-side wait. */
__global__ void nwb_inbox_ack_kernel(unsigned *ack, unsigned value)
{
    if (threadIdx.x == 0) nwb_st_release_sys(reinterpret_cast<int *>(ack), (int)value);
}
/* The inbox copy is zeroed by a kernel, not by cudaMemsetAsync: a memset queued behind a running fill sits in a
 * copy-engine queue until that fill is done, and the memsets / the summary upload with which the NEXT fills of the
 * other plans begin queue up behind it (measured: with five plans per GPU the fills ran in waves of four). */
__global__ void __launch_bounds__(128) nwb_inbox_zero_kernel(uint4 *dst, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = make_uint4(0u, 0u, 0u, 0u);
}
__global__ void __launch_bounds__(32) nwb_inbox_gate_kernel(const unsigned *peer_ack, unsigned need, int *err,
                                                           unsigned long long limit_ns)
{
    NwbWatchdog wd;
    for (;;) {
        const unsigned v = nwb_ld_relaxed_u32(peer_ack, true);
        if (__all_sync(NWB_FULL_MASK, (int)(v - need) >= 0)) break;
        if (wd.tick(err, limit_ns)) return;
        __nanosleep(200);
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
}

static int plan_run_body(nwb_plan *p, int m, int k, int d, void *stream, bool pipelined);

extern "C" int nwb_plan_run(nwb_plan *p, int m, int k, int d, void *stream)
{
    if (!p) return NWB_ERR_INVALID;
    return plan_run_body(p, m, k, d, stream, false);
}

extern "C" int nwb_plan_run_pipelined(nwb_plan *p, int m, int k, int d, void *stream)
{
    if (!p) return NWB_ERR_INVALID;
    int rc = plan_run_body(p, m, k, d, stream, true);
    if (rc != NWB_OK) return rc;
    if (p->inbox.base) {
        /* my copy of the inbox is free again: zero it, then tell the left neighbour */
        cudaStream_t st = p->last_stream;
        nwb_inbox_zero_kernel<<<64, 128, 0, st>>>(reinterpret_cast<uint4 *>(p->inbox.base + (size_t)(p->epoch & 1) * p->inbox.bytes),
                                                 p->inbox.bytes / 16);
        CK(cudaGetLastError());
        nwb_inbox_ack_kernel<<<1, 32, 0, st>>>(NWB_INBOX_ACK(p->inbox.base, p->inbox), (unsigned)(p->epoch + 1));
        CK(cudaGetLastError());
        p->launches += 1;
    }
    p->epoch++;
    return NWB_OK;
}

static int plan_run_body(nwb_plan *p, int m, int k, int d, void *stream, bool pipelined)
{
    CK(cudaSetDevice(p->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : p->stream;
    p->last_stream = st;
    p->m = m; p->k = k; p->d = d;
    const int A = p->A, B = p->B;
    unsigned flags = p->flags;
    if (flags & NWB_WANT_COUNT_MATRIX) flags |= NWB_WANT_COUNT;
    if (flags & NWB_WANT_COUNT_DIGEST) flags |= NWB_WANT_COUNT;
    if (p->suppress_count && !(flags & (NWB_WANT_COUNT_MATRIX | NWB_WANT_COUNT_DIGEST))) flags &= ~(unsigned)NWB_WANT_COUNT;
    const NwbTune tn = g_tune;

    NwbPkConsts pc;
    memset(&pc, 0, sizeof(pc));
    p->kind = choose_kind(flags, m, k, d, &pc);
    /* The count behind -s.  Default: the sparse backward sweep over the finished arrow codes
     * (nwb_count_sparse.cuh) with the dense forward sweep (nwb_count.cuh) launched behind it, which returns at
     * once unless the sparse sweep gave up.  A strip group (world > 1) keeps its arrow columns on different
     * GPUs and runs the dense sweep, whose strips hand their counts on like the fill's; the count matrix comes
     * from the general kernel's fused count. */
    int cpath = NWB_CNT_NONE;
    if (flags & NWB_WANT_COUNT) {
        if (flags & NWB_WANT_COUNT_MATRIX) cpath = NWB_CNT_FUSED;
        else if (tn.count_mode == 1) cpath = NWB_CNT_FUSED;
        else if (p->kind == NWB_KIND_I32 && p->world > 1) cpath = NWB_CNT_FUSED;
        else if (tn.count_mode >= 2 || (flags & NWB_WANT_COUNT_DIGEST) || p->world > 1) cpath = NWB_CNT_DENSE;
        else cpath = NWB_CNT_SPARSE;
    }
    if ((flags & NWB_WANT_COUNT_DIGEST) && cpath != NWB_CNT_DENSE) return NWB_ERR_UNSUPPORTED;
    p->count_path = cpath;
    p->count_pass = (cpath == NWB_CNT_DENSE || cpath == NWB_CNT_SPARSE);
    const bool fused_count = (cpath == NWB_CNT_FUSED);
    /* the kernels take the count request from the flags they are instantiated with */
    unsigned kflags = flags;
    if (!fused_count) kflags &= ~(unsigned)NWB_WANT_COUNT;
    int strip_w = NWB_I32_STRIP_W, pk_k = 0, pk_r = 1;
    p->pk_hx = false;
    p->pk_hy = false;
    p->pk_hz = false;
    if (p->kind == NWB_KIND_PK) {
        pk_k = (tn.pk_k == 1 || tn.pk_k == 2 || tn.pk_k == 4) ? tn.pk_k : nwb_pk_choose_k(A, B, p->world);
        if (p->count_pass) pk_k = 4; /* the count sweeps walk 256-column strips */
        /* two rows per step once the table is tall enough to amortise the doubled lane skew */
        pk_r = (tn.pk_r == 1 || tn.pk_r == 2) ? tn.pk_r : ((B >= 4096) ? 2 : 1);
        /* sweeping + flush warps (nwb_fill_hx.cuh) when every difference fits a nibble */
        const bool hx_ok = !fused_count && nwb_hx_supported(pc);
        p->pk_hx = hx_ok && pk_k == 4 && pk_r == 2;
        if (tn.pk_hx == 0) p->pk_hx = false;
        if (tn.pk_hx == 1 && hx_ok) { p->pk_hx = true; pk_k = 4; pk_r = 2; }
#ifdef NWB_EXPERIMENTS
        p->pk_hy = p->pk_hx && tn.pk_hy != 0;
#endif
        strip_w = 64 * pk_k;
    }
    p->L = nwb_make_layout(A, B, p->kind, pk_k, strip_w);
    p->L.pk_r = pk_r;
    const NwbLayout &L = p->L;

    /* this rank's strips */
    nwb_rank_strip_range(L.n_strips, p->rank, p->world, &p->strip_begin, &p->strip_end);
    const int nloc = p->strip_end - p->strip_begin;

    NwbDevSummary init;
    memset(&init, 0, sizeof(init));
    init.kernel_kind = p->kind;
    if (A == 0 || B == 0) {
        /* no interior cells: borders only (computation.c:97-124) */
        init.opt_score = (A == 0) ? -B * d : -A * d;
        init.count = (flags & NWB_WANT_COUNT) ? 1ull : 0ull;
    }
    CK(cudaMemcpyAsync(p->summary.p, &init, sizeof(init), cudaMemcpyHostToDevice, st));
    p->ran = true;
    p->timed = false;
    if (A == 0 || B == 0 || nloc <= 0) return NWB_OK;

    int rc = p->arrows.ensure(L.pitch * (size_t)B);
    if (rc == NWB_OK && (flags & NWB_WANT_SCORES)) rc = p->scores.ensure(L.spitch * (size_t)B);
    if (rc == NWB_OK && (flags & NWB_WANT_COUNT_MATRIX)) rc = p->cntmat.ensure(L.spitch * (size_t)B);
    if (rc == NWB_OK) rc = p->progress.ensure((size_t)nloc);
    if (rc == NWB_OK && p->kind == NWB_KIND_I32) {
        rc = p->bnd_s.ensure((size_t)nloc * L.bpitch);
        if (rc == NWB_OK && fused_count) rc = p->bnd_c.ensure((size_t)nloc * 2 * L.bpitch);
    }
    if (rc == NWB_OK && p->kind == NWB_KIND_PK) {
        rc = p->bnd_w.ensure((size_t)nloc * L.bpitch);
        if (rc == NWB_OK) rc = p->side_pre.ensure(NWB_PK_SPRE_LEN(B));
        if (rc == NWB_OK && fused_count) rc = p->bnd_c.ensure((size_t)nloc * 2 * L.bpitch);
    }
    /* the dense count sweep's own strips: 32 * cpl columns wide, aligned with this rank's 256-column fill strips */
    NwbCountParams cp;
    memset(&cp, 0, sizeof(cp));
    int cnt_cpl = 8;
    size_t cnt_stream_words = 0; /* uint64 words of this rank's count streams */
    if (rc == NWB_OK && p->count_pass) {
        long long cols = (long long)nloc * 256;
        if (cols > A) cols = A;
        cnt_cpl = (tn.cnt_cpl == 2 || tn.cnt_cpl == 4 || tn.cnt_cpl == 8) ? tn.cnt_cpl : nwb_count_choose_cpl(cols, p->sm_count);
        if (flags & NWB_WANT_COUNT_DIGEST) cnt_cpl = 8;
        const int wc = 32 * cnt_cpl, ratio = 256 / wc;
        const int sw_ratio = strip_w / 256 > 0 ? 1 : 256 / strip_w; /* fill strips per 256 columns (pk_k < 4 never has count_pass) */
        (void)sw_ratio;
        cp.n_strips = (A + wc - 1) / wc;
        cp.strip_begin = p->strip_begin * ratio < cp.n_strips ? p->strip_begin * ratio : cp.n_strips;
        cp.strip_end = p->strip_end * ratio < cp.n_strips ? p->strip_end * ratio : cp.n_strips;
        cnt_stream_words = (size_t)(cp.strip_end - cp.strip_begin) * 2 * L.bpitch;
        rc = p->bnd_c.ensure(cnt_stream_words);
    }
    if (rc != NWB_OK) return rc;
    CK(cudaMemsetAsync(p->progress.p, 0, (size_t)nloc * sizeof(int), st));
    /* the packed kernel's stream words validate themselves (bit 31): start from zero */
    if (p->kind == NWB_KIND_PK)
        CK(cudaMemsetAsync(p->bnd_w.p, 0, (size_t)nloc * L.bpitch * sizeof(uint32_t), st));
    if (p->kind == NWB_KIND_PK && fused_count)
        CK(cudaMemsetAsync(p->bnd_c.p, 0, (size_t)nloc * 2 * L.bpitch * sizeof(unsigned long long), st));
    if (cpath == NWB_CNT_DENSE)
        CK(cudaMemsetAsync(p->bnd_c.p, 0, cnt_stream_words * sizeof(unsigned long long), st));

    NwbStripParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.top = p->top.p;
    sp.side = p->side.p;
    sp.side_pre = p->side_pre.p;
    sp.A = A; sp.B = B; sp.m = m; sp.k = k; sp.d = d;
    sp.n_strips = L.n_strips;
    sp.strip_begin = p->strip_begin;
    sp.strip_end = p->strip_end;
    sp.arrows = p->arrows.p;
    sp.pitch = L.pitch;
    sp.scores = p->scores.p;
    sp.cntmat = p->cntmat.p;
    sp.spitch = L.spitch;
    sp.bnd_s = p->bnd_s.p;
    sp.bnd_c = p->bnd_c.p;
    sp.bnd_w = p->bnd_w.p;
    sp.bpitch = L.bpitch;
    sp.progress = p->progress.p;
    sp.summary = p->summary.p;
    /* branch counter (walk-table.c:108-120): the hx kernel's flush warps count for free; the one-warp-per-strip
     * packed kernel counts in its ring flush only on short tables -- on tall ones the extra instructions sit on
     * the strip-to-strip critical path (+1.6 ms at 100k x 100k) and a stand-alone HBM-bound pass (0.86 ms) is cheaper */
    const bool want_branches = !(flags & NWB_NO_BRANCH_COUNT);
    const bool branch_pass = want_branches && p->kind == NWB_KIND_PK && !p->pk_hx && B >= 4096;
    sp.count_branches = (want_branches && !branch_pass) ? 1 : 0;
    sp.debug_nowait = tn.inject_fault ? 4 : 0; /* test only: the watchdog turns the lost stream into NWB_ERR_CUDA */
#ifdef NWB_EXPERIMENTS
    sp.debug_nowait |= tn.debug_nowait;
#endif
    sp.watchdog_ns = (unsigned long long)(tn.watchdog_ms > 0 ? tn.watchdog_ms : 4000) * 1000000ull;
    /* a pipelined run e works with copy e & 1 of the inboxes (its own and the right neighbour's) */
    const size_t copy = pipelined ? (size_t)(p->epoch & 1) : 0;
    if (p->strip_begin > 0) {
        if (!p->inbox.base || L.bpitch > p->inbox.bpitch) return NWB_ERR_INVALID;
        const unsigned char *ib = p->inbox.base + copy * p->inbox.bytes;
        sp.in_bnd_s = (const int32_t *)(ib + p->inbox.off_s);
        sp.in_bnd_c = (const unsigned long long *)(ib + p->inbox.off_c);
        sp.in_bnd_w = (const uint32_t *)(ib + p->inbox.off_w);
        sp.in_progress = (const int *)(ib + p->inbox.off_flag);
    }
    if (p->strip_end < L.n_strips) {
        if (!p->right_base || L.bpitch > p->right.bpitch) return NWB_ERR_INVALID;
        unsigned char *ob = p->right_base + copy * p->right.bytes;
        sp.out_bnd_s = (int32_t *)(ob + p->right.off_s);
        sp.out_bnd_c = (unsigned long long *)(ob + p->right.off_c);
        sp.out_bnd_w = (uint32_t *)(ob + p->right.off_w);
        sp.out_progress = (int *)(ob + p->right.off_flag);
    }
    /* the copy of the neighbour's inbox this run writes into carried fill e - 2: the neighbour must have finished with
     * it and zeroed it.  The hx kernel in queue mode waits for that itself (its last strip, before its first remote
     * store); every other kernel gets the wait as a small kernel in front of it. */
    const bool need_gate = pipelined && p->strip_end < L.n_strips && p->epoch >= 2;

    int grid = nloc < p->sm_count ? nloc : p->sm_count;
    const bool hx = (p->kind == NWB_KIND_PK) && p->pk_hx;
    int hx_grid = grid;
    if (hx && (pipelined || (p->flags & NWB_QUEUE) || tn.hx_spb > 0)) {
        /* queue mode (nwb_fill_hx.cuh): blocks draw tickets and sweep three adjacent strips each -- as many blocks as
         * that takes, resident or not, so that the fills of other plans share the GPU with this one */
        int spb = (tn.hx_spb >= 1 && tn.hx_spb <= NWB_HX_CRIT) ? tn.hx_spb : NWB_HX_CRIT;
        sp.hx_spb = spb;
        hx_grid = (nloc + spb - 1) / spb;
        if (need_gate) {
            sp.gate_ack = NWB_INBOX_ACK(p->right_base, p->right);
            sp.gate_need = (unsigned)(p->epoch - 1);
        }
    }
    if (need_gate && !sp.gate_ack) {
        nwb_inbox_gate_kernel<<<1, 32, 0, st>>>(NWB_INBOX_ACK(p->right_base, p->right), (unsigned)(p->epoch - 1),
                                               &p->summary.p->error, sp.watchdog_ns);
        CK(cudaGetLastError());
        p->launches += 1;
    }
#ifdef NWB_EXPERIMENTS
    /* every strip of this launch can have an SM half (sweeping warp + packing warp + flush warp) to itself */
    p->pk_hz = hx && !p->pk_hy && tn.pk_hz != 0 && nwb_hz_usable(nloc, p->sm_count);
#endif
    /* one warp per SM sub-partition; a second one when there are more strips than that */
    int pk_warps = (nloc > p->sm_count * NWB_PK_WARPS) ? NWB_PK_MAX_WARPS : NWB_PK_WARPS;
    if (tn.pk_warps >= 1 && tn.pk_warps <= NWB_PK_MAX_WARPS) pk_warps = tn.pk_warps;
    if (p->kind == NWB_KIND_PK)
        while (pk_warps > 1 && NWB_PK_SMEM_BYTES(L.pk_k, L.pk_r, pk_warps) > 200 * 1024) pk_warps--;
    if (p->kind == NWB_KIND_PK && fused_count && pk_warps > NWB_PK_WARPS) pk_warps = NWB_PK_WARPS;
    CK(cudaEventRecord(p->ev0, st));
    if (p->kind == NWB_KIND_PK) {
        nwb_pk_prep_side_kernel<<<64, 256, 0, st>>>(p->side.p, B, pc.shift, p->side_pre.p);
        CK(cudaGetLastError());
        p->launches += 1;
    }
    /* count_mode 3 (opt-in): the dense count sweep trails the hx fill on a second stream: the flush warps
     * publish how many rows of each strip are in memory, the sweep waits for the rows it is about to read.  It
     * needs SMs the fill does not occupy (a fill block holds nearly all registers and shared memory of its SM)
     * and an otherwise idle GPU: the two kernels are separate cooperative launches, so their co-residency is
     * not guaranteed by CUDA -- the sweep's waits are bounded by the watchdog. */
    const int cnt_grid_wanted = (cp.strip_end - cp.strip_begin + NWB_CNT_WARPS - 1) / NWB_CNT_WARPS;
    const bool overlap = hx && cpath == NWB_CNT_DENSE && tn.count_mode == 3 && cnt_cpl == 8 && cp.strip_end > cp.strip_begin &&
                         !(flags & NWB_WANT_COUNT_DIGEST) && grid + cnt_grid_wanted <= p->sm_count && sp.hx_spb == 0;
    sp.publish_rows = overlap ? 1 : 0;
    if (overlap) CK(cudaEventRecord(p->ev_fork, st)); /* buffers are zeroed, strings uploaded */
    if (hx) {
#ifdef NWB_EXPERIMENTS
        rc = p->pk_hy ? nwb_hy_launch(sp, pc, grid, st, cuda_fail)
                      : (p->pk_hz ? nwb_hz_launch(sp, pc, grid, st, cuda_fail) : nwb_hx_launch(sp, pc, hx_grid, st, cuda_fail));
#else
        rc = nwb_hx_launch(sp, pc, hx_grid, st, cuda_fail);
#endif
    } else if (p->kind == NWB_KIND_PK) rc = run_pk(p, sp, pc, fused_count, grid, pk_warps, st);
    else rc = run_i32(p, kflags, sp, grid, st);
    if (rc != NWB_OK) return rc;
    if (branch_pass) {
        long long cb = (long long)p->strip_begin * L.strip_w, ce = (long long)p->strip_end * L.strip_w;
        if (ce > A) ce = A;
        nwb_branch_count_kernel<<<p->sm_count * 16, 256, 0, st>>>(p->arrows.p, L.pitch, A, B, (int)cb, (int)ce,
                                                                 &p->summary.p->branch_count);
        CK(cudaGetLastError());
        p->launches += 1;
    }
    if (cpath == NWB_CNT_SPARSE) {
        NwbSparseCountParams sc;
        memset(&sc, 0, sizeof(sc));
        sc.arrows = p->arrows.p;
        sc.pitch = L.pitch;
        sc.A = A; sc.B = B;
        sc.out_count = &p->summary.p->count;
        sc.out_state = &p->summary.p->count_state;
        sc.out_rows = &p->summary.p->sparse_rows;
        sc.mode = 0;
        nwb_sparse_count_kernel<<<1, 32, 0, st>>>(sc);
        CK(cudaGetLastError());
        /* the dense sweep's streams are zeroed only if it is going to run */
        nwb_zero_unless_done_kernel<<<p->sm_count * 4, 256, 0, st>>>(&p->summary.p->count_state, (uint4 *)p->bnd_c.p,
                                                                    cnt_stream_words * sizeof(unsigned long long) / 16);
        CK(cudaGetLastError());
        p->launches += 2;
    }
    if (p->count_pass && cp.strip_end > cp.strip_begin) {
        cp.arrows = p->arrows.p;
        cp.pitch = L.pitch;
        cp.A = A; cp.B = B;
        cp.bnd_c = p->bnd_c.p;
        cp.bpitch = L.bpitch;
        cp.in_bnd_c = sp.in_bnd_c;
        cp.out_bnd_c = sp.out_bnd_c;
        cp.summary = p->summary.p;
        cp.debug_nowait = sp.debug_nowait & 1;
        cp.watchdog_ns = sp.watchdog_ns;
        cp.fill_progress = overlap ? p->progress.p : nullptr;
        cp.skip_state = (cpath == NWB_CNT_SPARSE) ? &p->summary.p->count_state : nullptr;
        const int nlocc = cp.strip_end - cp.strip_begin;
        /* one warp per SM sub-partition on as few SMs as that takes */
        int cgrid = (nlocc + NWB_CNT_WARPS - 1) / NWB_CNT_WARPS;
        if (cgrid > p->sm_count) cgrid = p->sm_count;
        cudaStream_t cst = st;
        if (overlap) {
            cst = p->stream2;
            CK(cudaStreamWaitEvent(cst, p->ev_fork, 0));
        }
        rc = nwb_count_launch(cp, cnt_cpl, (flags & NWB_WANT_COUNT_DIGEST) != 0, cgrid, cst, cuda_fail);
        if (rc != NWB_OK) return rc;
        p->launches += 1;
        if (overlap) {
            CK(cudaEventRecord(p->ev_join, cst));
            CK(cudaStreamWaitEvent(st, p->ev_join, 0));
        }
    }
    CK(cudaEventRecord(p->ev1, st));
    p->timed = true;
    p->launches += 1;
    if (p->kind == NWB_KIND_PK && L.n_strips >= 2 && p->strip_end == L.n_strips) {
        /* this rank owns the last strip: add the score share carried by the stream it consumed */
        const uint32_t *stream = (p->strip_begin == L.n_strips - 1)
                                     ? sp.in_bnd_w
                                     : p->bnd_w.p + (size_t)(L.n_strips - 2 - p->strip_begin) * L.bpitch;
        nwb_pk_stream_sum_kernel<<<32, 256, 0, st>>>(stream, B, L.pk_r, &p->summary.p->rsum);
        CK(cudaGetLastError());
        p->launches += 1;
    }
    return NWB_OK;
}

static int run_pk(nwb_plan *p, const NwbStripParams &sp, const NwbPkConsts &pc, bool count, int grid, int warps,
                  cudaStream_t st)
{
    return nwb_pk_launch(sp, pc, p->L.pk_k, p->L.pk_r, count, grid, warps, st, cuda_fail);
}

extern "C" int nwb_plan_reset_inbox(nwb_plan *p, void *stream)
{
    if (!p) return NWB_ERR_INVALID;
    if (!p->inbox.base) return NWB_OK;
    CK(cudaSetDevice(p->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : p->stream;
    CK(cudaMemsetAsync(p->inbox.base, 0, NWB_INBOX_ALLOC_BYTES(p->inbox), st));
    p->epoch = 0; /* both copies and the acknowledgement word start over (every rank of the group does this) */
    return NWB_OK;
}

extern "C" int nwb_plan_summary(nwb_plan *p, nwb_summary *out)
{
    if (!p || !out) return NWB_ERR_INVALID;
    if (!p->ran) return NWB_ERR_INVALID;
    CK(cudaSetDevice(p->device));
    CK(cudaMemcpyAsync(&p->last, p->summary.p, sizeof(NwbDevSummary), cudaMemcpyDeviceToHost, p->last_stream));
    CK(cudaStreamSynchronize(p->last_stream));
    out->opt_score = p->last.opt_score;
    out->partial_r = 0;
    if (p->last.kernel_kind == NWB_KIND_PK && p->A > 0 && p->B > 0) {
        /* packed kernel: score(A,B) = sum_i u(i,B) - d*(A+B); a strip group sums the ranks' shares */
        out->partial_r = p->last.rsum;
        out->opt_score = (int32_t)(uint32_t)((unsigned long long)p->last.rsum -
                                             (unsigned long long)((long long)p->d * ((long long)p->A + p->B)));
    }
    out->branch_count = p->last.branch_count;
    out->greatest_abs = p->last.greatest_abs;
    out->kernel_kind = p->last.kernel_kind;
    out->count = p->last.count;
    out->count_path = p->count_path;
    if (p->count_path == NWB_CNT_SPARSE && p->last.count_state != NWB_SPC_DONE && p->A > 0 && p->B > 0 &&
        p->strip_end > p->strip_begin)
        out->count_path = NWB_CNT_SPARSE_BAILED;
    out->count_rows = p->last.sparse_rows;
    out->lastrow_count_digest = p->last.dig_row;
    out->lastcol_count_digest = p->last.dig_col;
    if (p->last.error != 0) {
        snprintf(g_cuda_err, sizeof(g_cuda_err),
                 "device watchdog: a strip waited more than the watchdog time for its left neighbour (code %d)", p->last.error);
        return NWB_ERR_CUDA;
    }
    return NWB_OK;
}

/* Digest of this plan's share of the arrow table (include/nwb.h, nwb_digest.cuh): rank digests add up. */
extern "C" int nwb_plan_arrow_digest(nwb_plan *p, uint64_t *out)
{
    if (!p || !out || !p->ran) return NWB_ERR_INVALID;
    *out = 0;
    if (p->A == 0 || p->B == 0 || p->strip_end <= p->strip_begin) return NWB_OK;
    CK(cudaSetDevice(p->device));
    cudaStream_t st = p->last_stream ? p->last_stream : p->stream;
    int rc = p->digest.ensure(1);
    if (rc != NWB_OK) return rc;
    CK(cudaMemsetAsync(p->digest.p, 0, sizeof(unsigned long long), st));
    long long cb = (long long)p->strip_begin * p->L.strip_w, ce = (long long)p->strip_end * p->L.strip_w;
    if (ce > p->A) ce = p->A;
    const int wb = (int)(cb / 8), we = (int)((ce + 7) / 8);
    nwb_arrow_digest_kernel<<<p->sm_count * 8, 256, 0, st>>>(p->arrows.p, p->L.pitch, p->A, p->B, wb, we, p->digest.p);
    CK(cudaGetLastError());
    p->launches += 1;
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, p->digest.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *out = h;
    return NWB_OK;
}

extern "C" float nwb_plan_kernel_ms(nwb_plan *p)
{
    if (!p || !p->timed) return 0.f;
    cudaSetDevice(p->device);
    float ms = 0.f;
    if (cudaEventSynchronize(p->ev1) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, p->ev0, p->ev1) != cudaSuccess) return -1.f;
    return ms;
}

extern "C" void *nwb_plan_arrows_device(nwb_plan *p) { return p ? (void *)p->arrows.p : nullptr; }
extern "C" size_t nwb_plan_arrow_pitch(const nwb_plan *p) { return p ? p->L.pitch : 0; }
extern "C" int64_t nwb_plan_launches(const nwb_plan *p) { return p ? p->launches : 0; }
extern "C" const char *nwb_plan_kernel_name(const nwb_plan *p)
{
    if (!p || !p->ran) return "";
    if (p->kind == NWB_KIND_I32) return "nwb_fill_i32_kernel";
    return p->pk_hx ? (p->pk_hy ? "nwb_fill_hy_kernel" : (p->pk_hz ? "nwb_fill_hz_kernel" : "nwb_fill_hx_kernel")) : "nwb_fill_pk_kernel";
}
/* How the last run obtained the count: "" (none), "fused", "dense" (forward sweep), "sparse" (backward sweep;
 * nwb_plan_summary() reports whether it had to fall back). */
extern "C" const char *nwb_plan_count_path_name(const nwb_plan *p)
{
    if (!p || !p->ran) return "";
    switch (p->count_path) {
    case NWB_CNT_FUSED: return "fused";
    case NWB_CNT_DENSE: return "dense";
    case NWB_CNT_SPARSE: return "sparse";
    default: return "";
    }
}
/* Host-only partition helpers (no device needed): what a launcher with one process per GPU uses to
 * shard the work the same way nwb_plan_create() / nwb_fill_on() do. */
extern "C" int nwb_strip_partition(int top_len, int strip_width, int rank, int world, int *begin_col, int *end_col)
{
    if (top_len < 0 || strip_width < 1 || world < 1 || rank < 0 || rank >= world) return NWB_ERR_INVALID;
    int n = (top_len + strip_width - 1) / strip_width;
    if (n < 1) n = 1;
    int sb, se;
    nwb_rank_strip_range(n, rank, world, &sb, &se);
    long long b = (long long)sb * strip_width, e = (long long)se * strip_width;
    if (b > top_len) b = top_len;
    if (e > top_len) e = top_len;
    if (begin_col) *begin_col = (int)b;
    if (end_col) *end_col = (int)e;
    return NWB_OK;
}
extern "C" int nwb_batch_partition(int64_t n_pairs, int rank, int world, int64_t *first_pair, int64_t *pair_count)
{
    if (n_pairs < 0 || world < 1 || rank < 0 || rank >= world) return NWB_ERR_INVALID;
    long long f, c;
    nwb_rank_pair_range(n_pairs, rank, world, &f, &c);
    if (first_pair) *first_pair = f;
    if (pair_count) *pair_count = c;
    return NWB_OK;
}
extern "C" int32_t nwb_strip_group_score(int64_t partial_r_sum, int top_len, int side_len, int d)
{
    return (int32_t)(uint32_t)((unsigned long long)partial_r_sum -
                               (unsigned long long)((long long)d * ((long long)top_len + side_len)));
}

extern "C" int nwb_plan_strip_range(const nwb_plan *p, int *begin_col, int *end_col)
{
    if (!p) return NWB_ERR_INVALID;
    long long b = (long long)p->strip_begin * p->L.strip_w, e = (long long)p->strip_end * p->L.strip_w;
    if (e > p->A) e = p->A;
    if (b > p->A) b = p->A;
    if (begin_col) *begin_col = (int)b;
    if (end_col) *end_col = (int)e;
    return NWB_OK;
}

extern "C" int nwb_plan_download_arrows(nwb_plan *p, uint8_t *dst, size_t dst_pitch, int row_begin, int row_end)
{
    if (!p || !dst || row_begin < 0 || row_end > p->B || row_begin > row_end) return NWB_ERR_INVALID;
    if (row_begin == row_end || p->A == 0) return NWB_OK;
    if (dst_pitch < (size_t)(p->A + 1) / 2) return NWB_ERR_INVALID;
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->last_stream ? p->last_stream : p->stream));
    const size_t width = dst_pitch < p->L.pitch ? dst_pitch : p->L.pitch;
    CK(cudaMemcpy2D(dst, dst_pitch, p->arrows.p + (size_t)row_begin * p->L.pitch, p->L.pitch, width,
                    (size_t)(row_end - row_begin), cudaMemcpyDeviceToHost));
    return NWB_OK;
}

/* ---- multi-process strips: CUDA IPC ---------------------------------------- */
extern "C" size_t nwb_plan_ipc_size(void) { return sizeof(NwbIpcBlob); }

extern "C" int nwb_plan_ipc_export(nwb_plan *p, void *blob)
{
    if (!p || !blob || !p->inbox.base) return NWB_ERR_INVALID;
    CK(cudaSetDevice(p->device));
    NwbIpcBlob b;
    memset(&b, 0, sizeof(b));
    CK(cudaIpcGetMemHandle(&b.handle, p->inbox.base));
    b.bpitch = p->inbox.bpitch; b.off_s = p->inbox.off_s; b.off_c = p->inbox.off_c;
    b.off_w = p->inbox.off_w; b.off_flag = p->inbox.off_flag; b.bytes = p->inbox.bytes;
    b.device = p->device;
    memcpy(blob, &b, sizeof(b));
    return NWB_OK;
}

extern "C" int nwb_plan_ipc_attach_right(nwb_plan *p, const void *blob)
{
    if (!p || !blob) return NWB_ERR_INVALID;
    CK(cudaSetDevice(p->device));
    NwbIpcBlob b;
    memcpy(&b, blob, sizeof(b));
    void *ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, b.handle, cudaIpcMemLazyEnablePeerAccess));
    p->right_base = (unsigned char *)ptr;
    p->right_is_ipc = true;
    p->right.bpitch = b.bpitch; p->right.off_s = b.off_s; p->right.off_c = b.off_c;
    p->right.off_w = b.off_w; p->right.off_flag = b.off_flag; p->right.bytes = b.bytes;
    p->right.base = p->right_base;
    return NWB_OK;
}

/* same-process variant: `right` lives on another device of this process */
static int plan_attach_right_local(nwb_plan *p, nwb_plan *right);
extern "C" int nwb_plan_attach_right(nwb_plan *p, nwb_plan *right)
{
    if (!p || !right || !right->inbox.base || p->device == right->device) return NWB_ERR_INVALID;
    return plan_attach_right_local(p, right);
}
static int plan_attach_right_local(nwb_plan *p, nwb_plan *right)
{
    CK(cudaSetDevice(p->device));
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, p->device, right->device));
    if (!can) return NWB_ERR_UNSUPPORTED;
    cudaError_t e = cudaDeviceEnablePeerAccess(right->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
    cudaGetLastError();
    p->right_base = right->inbox.base;
    p->right_is_ipc = false;
    p->right = right->inbox;
    return NWB_OK;
}

/* ========================================================================== */
struct nwb_table {
    int A = 0, B = 0;
    int m = 0, k = 0, d = 0;
    unsigned flags = 0;
    std::vector<nwb_plan *> plans;
    nwb_summary sum = {};
    float kernel_ms = 0.f;
    std::vector<char> top, side;
    uint8_t *h_arrows = nullptr;
    size_t pitch = 0;
    int32_t *h_scores = nullptr;
    unsigned long long *h_cntmat = nullptr;
    size_t spitch = 0;
    uint64_t arrow_digest = 0;
    bool have_digest = false;
};

extern "C" void nwb_free(nwb_table *t)
{
    if (!t) return;
    for (nwb_plan *p : t->plans) nwb_plan_destroy(p);
    free(t->h_arrows);
    free(t->h_scores);
    free(t->h_cntmat);
    delete t;
}

/* ---- workspace cache behind nwb_fill() / nwb_fill_on() --------------------------
 * A one-shot fill used to create and destroy its plans (streams, events, a dozen device buffers, their zeroing)
 * on every call: 3-4 ms of host time around a 0.75 ms kernel at 10k x 10k.  The plans of the last fills are kept
 * per (device, number of GPUs) and reused when they are large enough; nwb_cache_clear() releases them. */
struct NwbPlanSet {
    int device = 0, world = 1, maxA = 0, maxB = 0;
    bool busy = false, cached = false;
    std::vector<nwb_plan *> plans;
};
static std::mutex g_cache_mu;
static std::vector<NwbPlanSet *> g_cache;
#define NWB_CACHE_SETS 2

static void planset_destroy(NwbPlanSet *ps)
{
    if (!ps) return;
    for (nwb_plan *p : ps->plans) nwb_plan_destroy(p);
    delete ps;
}

static int planset_acquire(int device, int world, int A, int B, unsigned flags, NwbPlanSet **out)
{
    *out = nullptr;
    NwbPlanSet *ps = nullptr;
    std::vector<NwbPlanSet *> evict;
    const bool use_cache = g_tune.plan_cache != 0;
    if (use_cache) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (NwbPlanSet *c : g_cache)
            if (!c->busy && c->device == device && c->world == world && c->maxA >= A && c->maxB >= B) { ps = c; break; }
        if (ps) ps->busy = true;
        else {
            /* make room: drop idle sets of the same (device, world) that are too small, then the oldest idle one */
            for (size_t i = 0; i < g_cache.size();) {
                NwbPlanSet *c = g_cache[i];
                const bool same = c->device == device && c->world == world;
                if (!c->busy && (same || g_cache.size() >= NWB_CACHE_SETS)) {
                    evict.push_back(c);
                    g_cache.erase(g_cache.begin() + (long)i);
                } else i++;
            }
        }
    }
    for (NwbPlanSet *c : evict) planset_destroy(c);
    if (ps) {
        int rc = NWB_OK;
        for (nwb_plan *p : ps->plans) {
            p->flags = flags;
            if (world > 1 && rc == NWB_OK) rc = nwb_plan_reset_inbox(p, nullptr);
        }
        for (nwb_plan *p : ps->plans)
            if (world > 1 && rc == NWB_OK) {
                cudaSetDevice(p->device);
                cudaError_t e = cudaStreamSynchronize(p->stream);
                if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
            }
        if (rc != NWB_OK) {
            std::lock_guard<std::mutex> lk(g_cache_mu);
            ps->busy = false;
            return rc;
        }
        *out = ps;
        return NWB_OK;
    }
    ps = new (std::nothrow) NwbPlanSet();
    if (!ps) return NWB_ERR_NOMEM;
    ps->device = device;
    ps->world = world;
    ps->maxA = (int)nwb_round_up((size_t)(A > 4096 ? A : 4096), 1024);
    ps->maxB = (int)nwb_round_up((size_t)(B > 4096 ? B : 4096), 1024);
    ps->busy = true;
    int rc = NWB_OK;
    for (int g = 0; g < world && rc == NWB_OK; g++) {
        nwb_plan *p = nullptr;
        rc = nwb_plan_create(ps->maxA, ps->maxB, flags & ~(unsigned)(NWB_WANT_SCORES | NWB_WANT_COUNT_MATRIX), device + g, g, world, &p);
        if (rc == NWB_OK) {
            p->flags = flags;
            ps->plans.push_back(p);
        }
    }
    for (int g = 0; g + 1 < world && rc == NWB_OK; g++) rc = plan_attach_right_local(ps->plans[g], ps->plans[g + 1]);
    if (rc != NWB_OK) {
        planset_destroy(ps);
        return rc;
    }
    if (use_cache) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        if (g_cache.size() < NWB_CACHE_SETS) {
            g_cache.push_back(ps);
            ps->cached = true;
        }
    }
    *out = ps;
    return NWB_OK;
}

static void planset_release(NwbPlanSet *ps)
{
    if (!ps) return;
    if (ps->cached) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        ps->busy = false;
        return;
    }
    planset_destroy(ps);
}

extern "C" void nwb_cache_clear(void)
{
    std::vector<NwbPlanSet *> drop;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (size_t i = 0; i < g_cache.size();) {
            if (!g_cache[i]->busy) {
                drop.push_back(g_cache[i]);
                g_cache.erase(g_cache.begin() + (long)i);
            } else i++;
        }
    }
    for (NwbPlanSet *c : drop) planset_destroy(c);
}

/* The count of a strip group, tried on the rank that owns column A before anything else: the sparse backward sweep
 * (nwb_count_sparse.cuh) over that rank's own arrow columns.  On the BASELINE inputs the live cells die after ~1000
 * rows, long before the band reaches the rank's first column; if flow does get there the sweep gives up and the
 * caller runs the group's dense sweep.  Blocking. */
static int plan_group_sparse_count(nwb_plan *p, int *state, unsigned long long *count, unsigned *rows)
{
    CK(cudaSetDevice(p->device));
    cudaStream_t st = p->last_stream ? p->last_stream : p->stream;
    NwbSparseCountParams sc;
    memset(&sc, 0, sizeof(sc));
    sc.arrows = p->arrows.p;
    sc.pitch = p->L.pitch;
    sc.A = p->A; sc.B = p->B;
    sc.out_count = &p->summary.p->count;
    sc.out_state = &p->summary.p->count_state;
    sc.out_rows = &p->summary.p->sparse_rows;
    sc.min_col = p->strip_begin * p->L.strip_w + 1;
    nwb_sparse_count_kernel<<<1, 32, 0, st>>>(sc);
    CK(cudaGetLastError());
    p->launches += 1;
    NwbDevSummary h;
    CK(cudaMemcpyAsync(&h, p->summary.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *state = h.count_state;
    *count = h.count;
    *rows = h.sparse_rows;
    return NWB_OK;
}

extern "C" int nwb_fill_on(const char *top, int top_len, const char *side, int side_len,
                           int m, int k, int d, unsigned flags, int device, int num_gpus,
                           nwb_table **out)
{
    if (!out) return NWB_ERR_INVALID;
    *out = nullptr;
    if (top_len < 0 || side_len < 0 || (top_len && !top) || (side_len && !side) || num_gpus < 1)
        return NWB_ERR_INVALID;
    const int ndev = nwb_device_count();
    if (ndev <= 0) return NWB_ERR_NO_DEVICE;
    if (device < 0 || device + num_gpus > ndev) return NWB_ERR_INVALID;
    if (flags & NWB_WANT_COUNT_MATRIX) flags |= NWB_WANT_COUNT;
    if ((flags & (NWB_WANT_SCORES | NWB_WANT_COUNT_MATRIX)) && num_gpus > 1) return NWB_ERR_UNSUPPORTED;

    nwb_table *t = new (std::nothrow) nwb_table();
    if (!t) return NWB_ERR_NOMEM;
    t->A = top_len; t->B = side_len; t->m = m; t->k = k; t->d = d; t->flags = flags;
    t->top.assign(top, top + top_len);
    t->side.assign(side, side + side_len);

    NwbPlanSet *ps = nullptr;
    int rc = planset_acquire(device, num_gpus, top_len, side_len, flags, &ps);
    if (rc != NWB_OK) {
        nwb_free(t);
        return rc;
    }
    const std::vector<nwb_plan *> &plans = ps->plans;
    for (int g = 0; g < num_gpus && rc == NWB_OK; g++) rc = nwb_plan_upload(plans[g], top, top_len, side, side_len);
    /* -s on a strip group: fill first, then the sparse count on the rank that owns column A; only if that gives up
     * (flow reaches the rank's first column, band wider than the window) the group runs again with its dense sweep */
    const bool group_sparse = num_gpus > 1 && (flags & NWB_WANT_COUNT) && !(flags & (NWB_WANT_COUNT_MATRIX | NWB_WANT_COUNT_DIGEST)) &&
                              g_tune.count_mode == 0 && top_len > 0 && side_len > 0;
    int gs_state = NWB_SPC_NONE;
    unsigned long long gs_count = 0ull;
    unsigned gs_rows = 0u;
    for (nwb_plan *p : plans) p->suppress_count = group_sparse;
    for (int g = 0; g < num_gpus && rc == NWB_OK; g++) rc = nwb_plan_run(plans[g], m, k, d, nullptr);
    if (group_sparse && rc == NWB_OK) {
        nwb_plan *plast = nullptr;
        for (nwb_plan *p : plans)
            if (p->strip_end == p->L.n_strips && p->strip_end > p->strip_begin) plast = p;
        /* the fill of every rank must be over (and clean) before the arrow codes are read */
        for (int g = 0; g < num_gpus && rc == NWB_OK; g++) {
            nwb_summary s0;
            rc = nwb_plan_summary(plans[g], &s0);
        }
        if (rc == NWB_OK && plast) rc = plan_group_sparse_count(plast, &gs_state, &gs_count, &gs_rows);
        if (rc == NWB_OK && gs_state != NWB_SPC_DONE) {
            for (nwb_plan *p : plans) p->suppress_count = false;
            for (int g = 0; g < num_gpus && rc == NWB_OK; g++) rc = nwb_plan_reset_inbox(plans[g], nullptr);
            for (int g = 0; g < num_gpus && rc == NWB_OK; g++) {
                cudaSetDevice(plans[g]->device);
                cudaError_t e = cudaStreamSynchronize(plans[g]->stream);
                if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
            }
            for (int g = 0; g < num_gpus && rc == NWB_OK; g++) rc = nwb_plan_run(plans[g], m, k, d, nullptr);
        }
    }
    for (nwb_plan *p : plans) p->suppress_count = false;
    /* summaries: score/count live on the rank that owns column A (the last
     * non-empty one); branch counts and abs maxima are combined */
    memset(&t->sum, 0, sizeof(t->sum));
    for (int g = 0; g < num_gpus && rc == NWB_OK; g++) {
        nwb_summary s;
        rc = nwb_plan_summary(plans[g], &s);
        if (rc != NWB_OK) break;
        nwb_plan *p = plans[g];
        t->sum.branch_count += s.branch_count;
        if (s.greatest_abs > t->sum.greatest_abs) t->sum.greatest_abs = s.greatest_abs;
        t->sum.kernel_kind = s.kernel_kind;
        const bool owns_last = (top_len == 0 || side_len == 0) ? (g == 0)
                                                               : (p->strip_end == p->L.n_strips && p->strip_end > p->strip_begin);
        if (owns_last) {
            t->sum.opt_score = s.opt_score;
            t->sum.count = s.count;
            t->sum.count_path = s.count_path;
            t->sum.count_rows = s.count_rows;
            if (group_sparse) { /* which sweep delivered the group's count */
                if (gs_state == NWB_SPC_DONE) {
                    t->sum.count = gs_count;
                    t->sum.count_path = NWB_CNT_SPARSE;
                } else {
                    t->sum.count_path = NWB_CNT_SPARSE_BAILED;
                }
                t->sum.count_rows = gs_rows;
            }
        }
        t->sum.partial_r += s.partial_r;
        const float ms = nwb_plan_kernel_ms(p);
        if (ms > t->kernel_ms) t->kernel_ms = ms;
        t->sum.lastrow_count_digest += s.lastrow_count_digest;
        t->sum.lastcol_count_digest += s.lastcol_count_digest;
        if (flags & NWB_WANT_DIGEST) {
            uint64_t dg = 0;
            rc = nwb_plan_arrow_digest(p, &dg);
            if (rc != NWB_OK) break;
            t->arrow_digest += dg;
            t->have_digest = true;
        }
    }
    if (rc != NWB_OK) {
        /* a failed run (watchdog) may have left waiters on the other devices: let every stream drain */
        for (nwb_plan *p : plans) {
            cudaSetDevice(p->device);
            cudaStreamSynchronize(p->stream);
        }
    }
    if (rc == NWB_OK && top_len > 0 && side_len > 0 && t->sum.kernel_kind == NWB_KIND_PK)
        t->sum.opt_score = nwb_strip_group_score(t->sum.partial_r, top_len, side_len, d);
    if (rc == NWB_OK && top_len > 0 && side_len > 0) {
        const NwbLayout &L = plans[0]->L;
        t->pitch = L.pitch;
        t->spitch = L.spitch;
        if (flags & NWB_WANT_ARROWS_HOST) {
            t->h_arrows = (uint8_t *)malloc(L.pitch * (size_t)side_len);
            if (!t->h_arrows) rc = NWB_ERR_NOMEM;
            for (int g = 0; g < num_gpus && rc == NWB_OK; g++) {
                nwb_plan *p = plans[g];
                if (p->strip_end <= p->strip_begin) continue;
                const size_t off = (size_t)p->strip_begin * L.strip_w / 2;
                const size_t width = (size_t)(p->strip_end - p->strip_begin) * L.strip_w / 2;
                cudaSetDevice(p->device);
                cudaError_t e = cudaMemcpy2D(t->h_arrows + off, L.pitch, p->arrows.p + off, L.pitch, width,
                                             (size_t)side_len, cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy2D(arrows)");
            }
        }
        if (rc == NWB_OK && (flags & NWB_WANT_SCORES)) {
            t->h_scores = (int32_t *)malloc(L.spitch * (size_t)side_len * sizeof(int32_t));
            if (!t->h_scores) rc = NWB_ERR_NOMEM;
            else {
                cudaError_t e = cudaMemcpy(t->h_scores, plans[0]->scores.p,
                                           L.spitch * (size_t)side_len * sizeof(int32_t), cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy(scores)");
            }
        }
        if (rc == NWB_OK && (flags & NWB_WANT_COUNT_MATRIX)) {
            t->h_cntmat = (unsigned long long *)malloc(L.spitch * (size_t)side_len * sizeof(unsigned long long));
            if (!t->h_cntmat) rc = NWB_ERR_NOMEM;
            else {
                cudaError_t e = cudaMemcpy(t->h_cntmat, plans[0]->cntmat.p,
                                           L.spitch * (size_t)side_len * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy(cntmat)");
            }
        }
    }
    planset_release(ps);
    if (rc != NWB_OK) {
        nwb_free(t);
        return rc;
    }
    *out = t;
    return NWB_OK;
}

extern "C" int nwb_fill(const char *top, int top_len, const char *side, int side_len,
                        int m, int k, int d, unsigned flags, nwb_table **out)
{
    return nwb_fill_on(top, top_len, side, side_len, m, k, d, flags, 0, 1, out);
}

extern "C" int nwb_top_len(const nwb_table *t) { return t ? t->A : 0; }
extern "C" int nwb_side_len(const nwb_table *t) { return t ? t->B : 0; }
extern "C" int32_t nwb_opt_score(const nwb_table *t) { return t ? t->sum.opt_score : 0; }
extern "C" uint64_t nwb_count_u64(const nwb_table *t) { return t ? t->sum.count : 0; }
extern "C" uint32_t nwb_branch_count(const nwb_table *t) { return t ? t->sum.branch_count : 0; }
extern "C" int32_t nwb_greatest_abs_interior(const nwb_table *t) { return t ? t->sum.greatest_abs : 0; }
extern "C" float nwb_kernel_ms(const nwb_table *t) { return t ? t->kernel_ms : 0.f; }
extern "C" int nwb_kernel_kind(const nwb_table *t) { return t ? t->sum.kernel_kind : -1; }
extern "C" const uint8_t *nwb_arrow_rows(const nwb_table *t) { return t ? t->h_arrows : nullptr; }
extern "C" size_t nwb_arrow_pitch(const nwb_table *t) { return t ? t->pitch : 0; }
extern "C" int nwb_table_summary(const nwb_table *t, nwb_summary *out)
{
    if (!t || !out) return NWB_ERR_INVALID;
    *out = t->sum;
    return NWB_OK;
}
extern "C" int nwb_arrow_digest(const nwb_table *t, uint64_t *out)
{
    if (!t || !out) return NWB_ERR_INVALID;
    if (!t->have_digest && t->A > 0 && t->B > 0) return NWB_ERR_UNSUPPORTED; /* NWB_WANT_DIGEST was not set */
    *out = t->arrow_digest;
    return NWB_OK;
}

extern "C" int32_t nwb_score(const nwb_table *t, int i, int j)
{
    if (!t || i < 0 || j < 0 || i > t->A || j > t->B) return INT32_MIN;
    /* borders: init_computation_tables(), computation.c:97-124 */
    if (j == 0) return (int32_t)((uint32_t)i * (uint32_t)(-t->d));
    if (i == 0) return (int32_t)((uint32_t)j * (uint32_t)(-t->d));
    if (!t->h_scores) return INT32_MIN;
    return t->h_scores[(size_t)(j - 1) * t->spitch + (size_t)(i - 1)];
}

extern "C" unsigned nwb_arrows(const nwb_table *t, int i, int j)
{
    if (!t || i < 0 || j < 0 || i > t->A || j > t->B) return 0xFFFFFFFFu;
    if (i == 0 && j == 0) return 0;
    if (j == 0) return NWB_LEFT;
    if (i == 0) return NWB_UP;
    if (!t->h_arrows) return 0xFFFFFFFFu;
    const uint8_t byte = t->h_arrows[(size_t)(j - 1) * t->pitch + (size_t)((i - 1) >> 1)];
    unsigned code = ((i - 1) & 1) ? (byte >> 4) : (byte & 0xF);
    code &= 7u;
    if (t->top[(size_t)i - 1] == t->side[(size_t)j - 1]) code |= NWB_MATCH;
    return code;
}

extern "C" uint64_t nwb_count_at(const nwb_table *t, int i, int j)
{
    if (!t || i < 0 || j < 0 || i > t->A || j > t->B) return 0;
    if (i == 0 || j == 0) return 1;
    if (!t->h_cntmat) return 0;
    return t->h_cntmat[(size_t)(j - 1) * t->spitch + (size_t)(i - 1)];
}

extern "C" const int32_t *nwb_score_rows(const nwb_table *t, size_t *pitch_elems)
{
    if (pitch_elems) *pitch_elems = t ? t->spitch : 0;
    return t ? t->h_scores : nullptr;
}
extern "C" const uint64_t *nwb_count_rows(const nwb_table *t, size_t *pitch_elems)
{
    if (pitch_elems) *pitch_elems = t ? t->spitch : 0;
    return t ? (const uint64_t *)t->h_cntmat : nullptr;
}

/* ---- measurement aid: INT/DPX issue rate ---------------------- */
extern "C" int nwb_measure_int_issue(int device, int mode, double *per_clk_per_sm, double *gops_per_s)
{
    const int ndev = nwb_device_count();
    if (ndev <= 0) return NWB_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev || mode < 0 || mode > 3) return NWB_ERR_INVALID;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 2, threads = 1024, iters = 4096;
    int *sink = nullptr;
    long long *cyc = nullptr;
    CK(cudaMalloc((void **)&sink, sizeof(int)));
    CK(cudaMalloc((void **)&cyc, sizeof(long long) * blocks));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    long long max_cyc = 0;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(e0));
        switch (mode) {
        case 0: nwb_peak_kernel<0><<<blocks, threads>>>(iters, 3, -5, sink, cyc); break;
        case 1: nwb_peak_kernel<1><<<blocks, threads>>>(iters, 3, -5, sink, cyc); break;
        case 2: nwb_peak_kernel<2><<<blocks, threads>>>(iters, 0x00030004, 0x7fff7ffe, sink, cyc); break;
        default: nwb_peak_kernel<3><<<blocks, threads>>>(iters, 3, -5, sink, cyc); break;
        }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) {
            best_ms = ms;
            std::vector<long long> h(blocks);
            CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
            max_cyc = 0;
            for (long long c : h) if (c > max_cyc) max_cyc = c;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(cyc);
    /* thread-instructions: 64 per iteration per thread; two resident blocks share an SM */
    const double per_sm = 2.0 * threads * (double)iters * 64.0;
    if (per_clk_per_sm) *per_clk_per_sm = max_cyc > 0 ? per_sm / (double)max_cyc : 0.0;
    if (gops_per_s) *gops_per_s = (double)blocks * threads * (double)iters * 64.0 / (best_ms * 1e-3) / 1e9;
    return NWB_OK;
}

/* ========================================================================== */
#include "nwb_batch_api.inl"
