/* nwb_batch_api.inl -- host side of the batch entry points (include/nwb.h section 3). */
#define NWB_BATCH_MAX_CHUNKS 16
struct nwb_batch {
    int device = 0;
    unsigned flags = 0;
    int64_t n = 0;
    int m = 0, k = 0, d = 0;
    int sm_count = 0;
    int max_B = 0, max_strips = 1;
    long long max_A = 0;
    bool general = false; /* int32 engine, one warp per pair (nwb_batch_i32.cuh): any m/k/d, scores, |score| maximum */
    bool use_bx = false; /* two pairs per warp (nwb_batch_bx.cuh) */
    bool use_cx = false; /* ... swept back to back: every pair has the same shape */
    bool use_bp = false; /* bit-parallel rows, one thread per pair (nwb_batch_bp.cuh) */
    bool uniform = true;
    long long uni_A = -1, uni_B = -1;
    NwbPkConsts pc = {};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t copy_stream = nullptr;              /* nwb_batch_refill(): host-to-device copies of the next chunk */
    cudaStream_t stream2 = nullptr;                  /* ... odd chunks compute here: the kernels of neighbouring chunks overlap */
    cudaEvent_t ev_join = nullptr;
    cudaEvent_t ev_chunk[NWB_BATCH_MAX_CHUNKS] = {}; /* ... one "chunk is on the device" event per chunk */
    DevBuf<uint8_t> tops, sides, arrows;
    DevBuf<long long> top_off, side_off, arrow_off;
    DevBuf<int> score;
    DevBuf<unsigned> branch;
    DevBuf<uint32_t> scratch;
    DevBuf<unsigned long long> count, cscratch; /* NWB_WANT_COUNT */
    DevBuf<unsigned long long> digest;
    DevBuf<long long> fb_list; /* nwb_batch_bp_kernel: pairs left to nwb_batch_pk_kernel ... */
    DevBuf<unsigned> fb_count; /* ... and their number, one word per chunk of a refill */
    /* general (int32) path */
    DevBuf<int32_t> gscores, gbnd;
    DevBuf<long long> score_off;
    DevBuf<int> gabs, gprogress;
    DevBuf<NwbDevSummary> gsum;
    std::vector<long long> h_score_off;
    std::vector<int32_t> h_gscores;
    std::vector<int> h_abs;
    size_t gscores_elems = 0;
    std::vector<unsigned long long> h_count;
    std::vector<long long> h_top_off, h_side_off, h_arrow_off;
    std::vector<int> h_score;
    std::vector<unsigned> h_branch;
    std::vector<uint8_t> h_arrows;
    size_t arrows_bytes = 0;
    bool ran = false, fetched = false;
    int64_t launches = 0;
};

extern "C" void nwb_batch_free(nwb_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    b->tops.release(); b->sides.release(); b->arrows.release(); b->top_off.release(); b->side_off.release();
    b->arrow_off.release(); b->score.release(); b->branch.release(); b->scratch.release();
    b->count.release(); b->cscratch.release(); b->digest.release(); b->fb_list.release(); b->fb_count.release();
    b->gscores.release(); b->gbnd.release(); b->score_off.release(); b->gabs.release(); b->gprogress.release(); b->gsum.release();
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    for (int i = 0; i < NWB_BATCH_MAX_CHUNKS; i++)
        if (b->ev_chunk[i]) cudaEventDestroy(b->ev_chunk[i]);
    if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
    if (b->stream2) cudaStreamDestroy(b->stream2);
    if (b->ev_join) cudaEventDestroy(b->ev_join);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

static int batch_create_impl(const char *tops, const int64_t *top_off, const char *sides, const int64_t *side_off,
                             int64_t n_pairs, int m, int k, int d, unsigned flags, int device, bool upload_strings,
                             nwb_batch **out)
{
    if (!out) return NWB_ERR_INVALID;
    *out = nullptr;
    if (n_pairs < 0 || !top_off || !side_off) return NWB_ERR_INVALID;
    if (flags & NWB_WANT_COUNT_MATRIX) return NWB_ERR_UNSUPPORTED; /* per-cell counts: nwb_fill() per pair */
    const int ndev = nwb_device_count();
    if (ndev <= 0) return NWB_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return NWB_ERR_INVALID;
    NwbPkConsts pc;
    memset(&pc, 0, sizeof(pc));
    /* schemes outside the packed kernels' range (the reference takes whatever atoi() yields,
     * needleman-wunsch.c:783-785), the score matrix and the |score| maximum run on the int32 engine */
    const bool general = (flags & (NWB_WANT_SCORES | NWB_TRACK_ABS | NWB_FORCE_GENERAL)) != 0 || !nwb_pk_supported(m, k, d, &pc);
    CK(cudaSetDevice(device));
    nwb_batch *b = new (std::nothrow) nwb_batch();
    if (!b) return NWB_ERR_NOMEM;
    b->device = device; b->flags = flags; b->n = n_pairs; b->m = m; b->k = k; b->d = d; b->pc = pc;
    b->general = general;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev1);
    if (e != cudaSuccess) { int rc = cuda_fail(e, "batch setup"); nwb_batch_free(b); return rc; }
    b->sm_count = prop.multiProcessorCount;
    b->h_top_off.assign(top_off, top_off + n_pairs + 1);
    b->h_side_off.assign(side_off, side_off + n_pairs + 1);
    b->h_arrow_off.resize((size_t)n_pairs + 1);
    if (general && (flags & NWB_WANT_SCORES)) b->h_score_off.resize((size_t)n_pairs + 1);
    long long aoff = 0, soff = 0;
    for (int64_t p = 0; p < n_pairs; p++) {
        const long long A = top_off[p + 1] - top_off[p], B = side_off[p + 1] - side_off[p];
        if (A < 0 || B < 0 || A > INT_MAX / 2 || B > 60000) { nwb_batch_free(b); return NWB_ERR_INVALID; }
        const int ns = (int)((A + 255) / 256);
        if (ns > b->max_strips) b->max_strips = ns;
        if ((int)B > b->max_B) b->max_B = (int)B;
        if (A > b->max_A) b->max_A = A;
        if (p == 0) { b->uni_A = A; b->uni_B = B; }
        else if (A != b->uni_A || B != b->uni_B) b->uniform = false;
        b->h_arrow_off[(size_t)p] = aoff;
        aoff += (long long)(ns > 0 ? ns : 1) * 128 * B;
        if (!b->h_score_off.empty()) {
            b->h_score_off[(size_t)p] = soff;
            soff += (long long)(ns > 0 ? ns : 1) * 256 * B;
        }
    }
    b->h_arrow_off[(size_t)n_pairs] = aoff;
    b->arrows_bytes = (size_t)aoff;
    if (!b->h_score_off.empty()) b->h_score_off[(size_t)n_pairs] = soff;
    b->gscores_elems = (size_t)soff;
    if (!general) {
        /* short top strings and nibble-sized differences: two pairs per warp (nwb_tune "batch_bx" / "batch_cx" = 0
         * keep the simpler kernels, for tests) */
        b->use_bx = nwb_bx_usable(pc, b->max_A, b->max_B) && g_tune.batch_bx != 0;
        b->use_cx = b->use_bx && n_pairs > 0 && nwb_cx_usable(pc, b->uniform, b->uni_A, (int)b->uni_B) && g_tune.batch_cx != 0;
        /* small differences (2d + m <= 3: DNA 1/1/1) and one strip: a thread per pair, a row per addition (row vectors
         * of 64, 128 or 256 bits by the longest top string; very short strings keep the warp kernels) */
        b->use_bp = n_pairs > 0 && nwb_bp_usable(pc, b->max_A) && g_tune.batch_bp != 0 && (b->max_A > NWB_BP_MIN_AUTO || g_tune.batch_bp == 1);
    }
    if (!general && !b->use_bx && NWB_BATCH_SMEM_PER_WARP(b->max_B) > 220 * 1024) { nwb_batch_free(b); return NWB_ERR_UNSUPPORTED; }
    const size_t tbytes = (size_t)top_off[n_pairs], sbytes = (size_t)side_off[n_pairs];
    int rc = b->tops.ensure(tbytes + 16);
    if (rc == NWB_OK) rc = b->sides.ensure(sbytes + 16);
    if (rc == NWB_OK) rc = b->top_off.ensure((size_t)n_pairs + 1);
    if (rc == NWB_OK) rc = b->side_off.ensure((size_t)n_pairs + 1);
    if (rc == NWB_OK) rc = b->arrow_off.ensure((size_t)n_pairs + 1);
    if (rc == NWB_OK) rc = b->arrows.ensure(b->arrows_bytes + 16);
    if (rc == NWB_OK) rc = b->score.ensure((size_t)n_pairs + 1);
    if (rc == NWB_OK) rc = b->branch.ensure((size_t)n_pairs + 1);
    if (rc == NWB_OK && (flags & NWB_WANT_COUNT)) rc = b->count.ensure((size_t)n_pairs + 1);
    if (rc == NWB_OK && (b->use_bp || (flags & NWB_WANT_COUNT))) rc = b->fb_list.ensure((size_t)n_pairs + 1);
    if (rc == NWB_OK && (b->use_bp || (flags & NWB_WANT_COUNT))) rc = b->fb_count.ensure(NWB_BATCH_MAX_CHUNKS + 1);
    if (rc == NWB_OK && general) {
        const size_t nwarps = (size_t)b->sm_count * 2 * NWB_BI32_WARPS;
        const size_t bpitch = nwb_round_up((size_t)b->max_B + 2, 32);
        rc = b->gbnd.ensure(nwarps * (size_t)b->max_strips * bpitch);
        if (rc == NWB_OK) rc = b->gprogress.ensure(nwarps * (size_t)b->max_strips);
        if (rc == NWB_OK) rc = b->gsum.ensure(nwarps);
        if (rc == NWB_OK && (flags & NWB_TRACK_ABS)) rc = b->gabs.ensure((size_t)n_pairs + 1);
        if (rc == NWB_OK && (flags & NWB_WANT_SCORES)) {
            rc = b->gscores.ensure(b->gscores_elems + 16);
            if (rc == NWB_OK) rc = b->score_off.ensure((size_t)n_pairs + 1);
        }
    }
    if (rc != NWB_OK) { nwb_batch_free(b); return rc; }
    e = cudaSuccess;
    if (upload_strings && tbytes) e = cudaMemcpyAsync(b->tops.p, tops, tbytes, cudaMemcpyHostToDevice, b->stream);
    if (upload_strings && e == cudaSuccess && sbytes) e = cudaMemcpyAsync(b->sides.p, sides, sbytes, cudaMemcpyHostToDevice, b->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b->top_off.p, b->h_top_off.data(), ((size_t)n_pairs + 1) * 8, cudaMemcpyHostToDevice, b->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b->side_off.p, b->h_side_off.data(), ((size_t)n_pairs + 1) * 8, cudaMemcpyHostToDevice, b->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b->arrow_off.p, b->h_arrow_off.data(), ((size_t)n_pairs + 1) * 8, cudaMemcpyHostToDevice, b->stream);
    if (e == cudaSuccess && !b->h_score_off.empty())
        e = cudaMemcpyAsync(b->score_off.p, b->h_score_off.data(), ((size_t)n_pairs + 1) * 8, cudaMemcpyHostToDevice, b->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
    if (e != cudaSuccess) { rc = cuda_fail(e, "batch upload"); nwb_batch_free(b); return rc; }
    *out = b;
    return NWB_OK;
}

extern "C" int nwb_batch_create(const char *tops, const int64_t *top_off, const char *sides, const int64_t *side_off,
                                int64_t n_pairs, int m, int k, int d, unsigned flags, int device, nwb_batch **out)
{
    if (n_pairs > 0 && (!tops || !sides) && top_off && side_off && (top_off[n_pairs] > 0 || side_off[n_pairs] > 0)) return NWB_ERR_INVALID;
    return batch_create_impl(tops, top_off, sides, side_off, n_pairs, m, k, d, flags, device, true, out);
}

/* the count behind -s: a second pass over the arrow codes the fill has just written (nwb_batch_count.cuh) */
static int batch_count_pass(nwb_batch *b, cudaStream_t st, int64_t c0, int64_t c1, int chunk)
{
    NwbBatchCountParams cp;
    memset(&cp, 0, sizeof(cp));
    const int grid = b->sm_count;
    cp.top_off = b->top_off.p + c0; cp.side_off = b->side_off.p + c0; cp.n_pairs = c1 - c0;
    cp.arrows = b->arrows.p; cp.arrow_off = b->arrow_off.p + c0; cp.out_count = b->count.p + c0;
    if (b->max_strips > 1) {
        cp.scratch_per_warp = nwb_round_up((size_t)b->max_B + 1, 16);
        int rc = b->cscratch.ensure((size_t)grid * NWB_BCNT_WARPS * cp.scratch_per_warp);
        if (rc != NWB_OK) return rc;
        cp.scratch = b->cscratch.p;
    }
    const size_t smem = (size_t)NWB_BCNT_SMEM_PER_WARP * NWB_BCNT_WARPS;
    if (g_tune.batch_lcount != 0 && b->fb_list.p && b->fb_count.p) {
        /* one thread per pair, backwards over the cells on optimal paths (nwb_batch_lcount.cuh); the pairs whose band
         * does not fit its window land on a list that the dense kernel works off right behind (the list of this chunk:
         * the fill's own left-over launch, earlier on this stream, is done with it) */
        unsigned *fbc = b->fb_count.p + chunk;
        CK(cudaMemsetAsync(fbc, 0, sizeof(unsigned), st));
        NwbLaneCountParams lp;
        memset(&lp, 0, sizeof(lp));
        lp.top_off = cp.top_off; lp.side_off = cp.side_off; lp.n_pairs = cp.n_pairs; lp.arrows = cp.arrows;
        lp.arrow_off = cp.arrow_off; lp.out_count = cp.out_count; lp.fb_list = b->fb_list.p + c0; lp.fb_count = fbc;
        int lw = g_tune.lc_warps > 0 ? g_tune.lc_warps : nwb_lc_choose_warps((cp.n_pairs + 31) / 32, grid);
        if (lw > NWB_LC_WARPS) lw = NWB_LC_WARPS;
        nwb_batch_lcount_kernel<<<grid, 32 * lw, 0, st>>>(lp);
        CK(cudaGetLastError());
        cp.pair_list = lp.fb_list; cp.pair_count = fbc;
        CK(cudaFuncSetAttribute(nwb_batch_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nwb_batch_count_kernel<<<grid, 32 * NWB_BCNT_WARPS, smem, st>>>(cp);
        CK(cudaGetLastError());
        b->launches += 2;
        return NWB_OK;
    }
    /* every pair the same one-strip shape: the tables of a warp's run of pairs are swept back to back
     * (nwb_tune "bcnt_chain" = 0: one pair at a time) */
    if (nwb_bcount_chain_usable(b->uniform, b->uni_A, b->uni_B, c1 - c0, (long long)grid * NWB_BCNT_WARPS) && g_tune.bcnt_chain != 0) {
        CK(cudaFuncSetAttribute(nwb_batch_count_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nwb_batch_count_chain_kernel<<<grid, 32 * NWB_BCNT_WARPS, smem, st>>>(cp, (int)b->uni_A, (int)b->uni_B);
        CK(cudaGetLastError());
        b->launches += 1;
        return NWB_OK;
    }
    CK(cudaFuncSetAttribute(nwb_batch_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nwb_batch_count_kernel<<<grid, 32 * NWB_BCNT_WARPS, smem, st>>>(cp);
    CK(cudaGetLastError());
    b->launches += 1;
    return NWB_OK;
}

static int batch_fill_pass(nwb_batch *b, cudaStream_t st, int64_t c0, int64_t c1, int chunk);

extern "C" int nwb_batch_run(nwb_batch *b, void *stream)
{
    if (!b) return NWB_ERR_INVALID;
    CK(cudaSetDevice(b->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : b->stream;
    b->ran = true;
    b->fetched = false;
    if (b->n == 0) return NWB_OK;
    CK(cudaEventRecord(b->ev0, st));
    int rc = batch_fill_pass(b, st, 0, b->n, 0);
    if (rc == NWB_OK && (b->flags & NWB_WANT_COUNT)) rc = batch_count_pass(b, st, 0, b->n, 0);
    if (rc != NWB_OK) return rc;
    CK(cudaEventRecord(b->ev1, st));
    return NWB_OK;
}

/* New strings for the same shapes, from HOST buffers: the batch is cut into chunks of pairs, the host-to-device
 * copy of chunk i+1 (copy stream) overlaps the kernels of chunk i (the batch's own stream).  tops/sides hold the
 * pairs' strings at the offsets given to nwb_batch_create(). */
extern "C" int nwb_batch_refill(nwb_batch *b, const char *tops, const char *sides)
{
    if (!b || (b->n > 0 && (!tops || !sides))) return NWB_ERR_INVALID;
    CK(cudaSetDevice(b->device));
    b->ran = true;
    b->fetched = false;
    if (b->n == 0) return NWB_OK;
    if (!b->copy_stream) {
        CK(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&b->stream2, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming));
        for (int i = 0; i < NWB_BATCH_MAX_CHUNKS; i++) CK(cudaEventCreateWithFlags(&b->ev_chunk[i], cudaEventDisableTiming));
    }
    cudaStream_t st = b->stream;
    int64_t nchunks = b->n / 16384;
    if (nchunks < 1) nchunks = 1;
    if (nchunks > 8) nchunks = 8;
    /* two pairs share a warp in the bx/cx kernels: keep chunk boundaries even */
    int64_t per = (b->n + nchunks - 1) / nchunks;
    per += per & 1;
    CK(cudaEventRecord(b->ev0, st));
    int rc = NWB_OK;
    int ci = 0;
    /* one-strip pairs on the packed / bit-parallel kernels keep everything per warp in shared memory; wider pairs
     * and the int32 engine have per-warp scratch in global memory that two launches in flight would share */
    const bool two_streams = !b->general && b->max_strips == 1;
    for (int64_t c0 = 0; c0 < b->n && rc == NWB_OK; c0 += per, ci++) {
        const int64_t c1 = (c0 + per < b->n) ? c0 + per : b->n;
        const long long tb = b->h_top_off[(size_t)c0], te = b->h_top_off[(size_t)c1];
        const long long sb = b->h_side_off[(size_t)c0], se = b->h_side_off[(size_t)c1];
        if (te > tb) CK(cudaMemcpyAsync(b->tops.p + tb, tops + tb, (size_t)(te - tb), cudaMemcpyHostToDevice, b->copy_stream));
        if (se > sb) CK(cudaMemcpyAsync(b->sides.p + sb, sides + sb, (size_t)(se - sb), cudaMemcpyHostToDevice, b->copy_stream));
        CK(cudaEventRecord(b->ev_chunk[ci], b->copy_stream));
        /* A chunk is a fraction of the batch and does not fill the GPU on its own (one thread per pair in
         * nwb_batch_bp_kernel: 16,000 pairs are 4 warps per SM): odd chunks go to a second stream so that the
         * kernels of neighbouring chunks run side by side. */
        cudaStream_t cst = ((ci & 1) && two_streams) ? b->stream2 : st;
        if (ci == 1 && two_streams) CK(cudaStreamWaitEvent(b->stream2, b->ev0, 0));
        CK(cudaStreamWaitEvent(cst, b->ev_chunk[ci], 0));
        rc = batch_fill_pass(b, cst, c0, c1, ci);
        if (rc == NWB_OK && (b->flags & NWB_WANT_COUNT)) rc = batch_count_pass(b, cst, c0, c1, ci);
    }
    if (rc != NWB_OK) return rc;
    if (ci > 1 && two_streams) {
        CK(cudaEventRecord(b->ev_join, b->stream2));
        CK(cudaStreamWaitEvent(st, b->ev_join, 0));
    }
    CK(cudaEventRecord(b->ev1, st));
    return NWB_OK;
}

static int batch_fill_pass(nwb_batch *b, cudaStream_t st, int64_t c0, int64_t c1, int chunk)
{
    if (b->general) {
        NwbBatchI32Params gp;
        memset(&gp, 0, sizeof(gp));
        gp.tops = b->tops.p; gp.top_off = b->top_off.p + c0; gp.sides = b->sides.p; gp.side_off = b->side_off.p + c0;
        gp.n_pairs = c1 - c0; gp.m = b->m; gp.k = b->k; gp.d = b->d;
        gp.arrows = b->arrows.p; gp.arrow_off = b->arrow_off.p + c0;
        gp.scores = b->gscores.p; gp.score_off = b->score_off.p ? b->score_off.p + c0 : nullptr;
        gp.out_score = b->score.p + c0;
        gp.out_branch = (b->flags & NWB_NO_BRANCH_COUNT) ? nullptr : b->branch.p + c0;
        gp.out_abs = (b->flags & NWB_TRACK_ABS) ? b->gabs.p + c0 : nullptr;
        gp.bnd_s = b->gbnd.p;
        gp.bpitch = nwb_round_up((size_t)b->max_B + 2, 32);
        gp.progress = b->gprogress.p;
        gp.max_strips = b->max_strips;
        gp.wsum = b->gsum.p;
        const int ggrid = b->sm_count * 2;
        const size_t smem = (size_t)NWB_BI32_WARPS * NWB_I32_STAGE_WORDS * 4;
        const bool S = (b->flags & NWB_WANT_SCORES) != 0, AB = (b->flags & NWB_TRACK_ABS) != 0;
#define NWB_BI32_GO(s_, a_)                                                                                          \
    do {                                                                                                             \
        CK(cudaFuncSetAttribute(nwb_batch_i32_kernel<s_, a_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        nwb_batch_i32_kernel<s_, a_><<<ggrid, 32 * NWB_BI32_WARPS, smem, st>>>(gp);                                    \
    } while (0)
        if (S) { if (AB) NWB_BI32_GO(true, true); else NWB_BI32_GO(true, false); }
        else { if (AB) NWB_BI32_GO(false, true); else NWB_BI32_GO(false, false); }
#undef NWB_BI32_GO
        CK(cudaGetLastError());
        b->launches += 1;
        return NWB_OK;
    }
    NwbBatchParams bp;
    memset(&bp, 0, sizeof(bp));
    const int grid = b->sm_count;
    if (b->use_bp) {
        /* one thread per pair; whatever it cannot take (more than five letters in a top string) lands on a list
         * that the one-warp-per-pair kernel works off right behind it (a launch that finds the list empty returns) */
        unsigned *fbc = b->fb_count.p + chunk; /* every chunk of a refill has its own list and counter */
        CK(cudaMemsetAsync(fbc, 0, sizeof(unsigned), st));
        NwbBpParams pp;
        memset(&pp, 0, sizeof(pp));
        pp.tops = b->tops.p; pp.top_off = b->top_off.p + c0; pp.sides = b->sides.p; pp.side_off = b->side_off.p + c0;
        pp.n_pairs = c1 - c0; pp.d = b->d;
        pp.arrows = b->arrows.p; pp.arrow_off = b->arrow_off.p + c0; pp.out_score = b->score.p + c0;
        pp.out_branch = (b->flags & NWB_NO_BRANCH_COUNT) ? nullptr : b->branch.p + c0;
        pp.fb_list = b->fb_list.p + c0; pp.fb_count = fbc; pp.k2 = 2u; pp.k4 = 4u;
        int warps = g_tune.bp_warps > 0 ? g_tune.bp_warps : nwb_bp_choose_warps((c1 - c0 + 31) / 32, grid);
        if (warps > NWB_BP_WARPS) warps = NWB_BP_WARPS;
        const bool aligned = warps >= NWB_BP_ALIGNED_MIN_WARPS && g_tune.bp_aligned != 0;
        const size_t smem = aligned ? (size_t)NWB_BP_SMEM_MAX : NWB_BP_SMEM_BYTES(warps);
        int launched = 0;
#define NWB_BP_GO_A(M_, N_, W_, A_)                                                                                        \
    {                                                                                                                       \
        CK(cudaFuncSetAttribute(nwb_batch_bp_kernel<M_, N_, W_, A_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        nwb_batch_bp_kernel<M_, N_, W_, A_><<<grid, 32 * warps, smem, st>>>(pp);                                               \
        launched = 1;                                                                                                       \
    }
#define NWB_BP_GO_W(M_, N_, W_) { if (aligned) NWB_BP_GO_A(M_, N_, W_, true) else NWB_BP_GO_A(M_, N_, W_, false) }
#define NWB_BP_GO(M_, N_)                                                                                            \
    if (b->pc.a_match == M_ && b->pc.a_mis == N_) {                                                                  \
        if (nw == 2) NWB_BP_GO_W(M_, N_, 2) else if (nw == 4) NWB_BP_GO_W(M_, N_, 4) else NWB_BP_GO_W(M_, N_, 8)        \
    }
        const int nw = nwb_bp_words(b->max_A); /* 64-, 128- or 256-bit row vectors */
        NWB_BP_GO(1, 0) NWB_BP_GO(1, 1) NWB_BP_GO(2, 0) NWB_BP_GO(2, 1) NWB_BP_GO(2, 2)
        NWB_BP_GO(3, 0) NWB_BP_GO(3, 1) NWB_BP_GO(3, 2) NWB_BP_GO(3, 3)
#undef NWB_BP_GO_W
#undef NWB_BP_GO_A
#undef NWB_BP_GO
        if (!launched) return NWB_ERR_UNSUPPORTED;
        CK(cudaGetLastError());
        int pw = (int)((220 * 1024) / NWB_BATCH_SMEM_PER_WARP(b->max_B));
        if (pw > NWB_BATCH_WARPS) pw = NWB_BATCH_WARPS;
        if (pw < 1) return NWB_ERR_UNSUPPORTED;
        bp.tops = b->tops.p; bp.top_off = b->top_off.p + c0; bp.sides = b->sides.p; bp.side_off = b->side_off.p + c0;
        bp.n_pairs = c1 - c0; bp.m = b->m; bp.k = b->k; bp.d = b->d; bp.max_B = b->max_B;
        bp.arrows = b->arrows.p; bp.arrow_off = b->arrow_off.p + c0; bp.out_score = b->score.p + c0;
        bp.out_branch = (b->flags & NWB_NO_BRANCH_COUNT) ? nullptr : b->branch.p + c0;
        bp.bpitch = nwb_round_up((size_t)b->max_B + 1 + 64 + 256, 32);
        bp.pair_list = pp.fb_list; bp.pair_count = fbc;
        const size_t psmem = NWB_BATCH_SMEM_PER_WARP(b->max_B) * (size_t)pw;
        CK(cudaFuncSetAttribute(nwb_batch_pk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        nwb_batch_pk_kernel<<<grid, 32 * pw, psmem, st>>>(bp, b->pc);
        CK(cudaGetLastError());
        b->launches += 2;
        return NWB_OK;
    }
    if (b->use_bx) {
        int warps = (int)((220 * 1024) / NWB_BX_SMEM_PER_WARP(b->max_B));
        if (warps > NWB_BX_WARPS) warps = NWB_BX_WARPS;
        if (warps < 1) return NWB_ERR_UNSUPPORTED;
        bp.tops = b->tops.p; bp.top_off = b->top_off.p + c0; bp.sides = b->sides.p; bp.side_off = b->side_off.p + c0;
        bp.n_pairs = c1 - c0; bp.m = b->m; bp.k = b->k; bp.d = b->d; bp.max_B = b->max_B;
        bp.arrows = b->arrows.p; bp.arrow_off = b->arrow_off.p + c0; bp.out_score = b->score.p + c0;
        bp.out_branch = (b->flags & NWB_NO_BRANCH_COUNT) ? nullptr : b->branch.p + c0;
        if (b->use_cx) {
            /* uniform shapes: the kernel addresses pair p's table as arrows + p * table bytes (not through arrow_off) */
            bp.arrows = b->arrows.p + b->h_arrow_off[(size_t)c0];
            const bool w16 = g_tune.cx_warps == 16 && nwb_cx_usable(b->pc, true, b->uni_A, (int)b->uni_B, 16);
            const int cw = w16 ? 16 : NWB_BX_WARPS;
            const size_t smem = NWB_CX_SMEM_PER_WARP(b->max_B, cw) * (size_t)cw;
            auto kern = w16 ? nwb_batch_cx_kernel<16> : nwb_batch_cx_kernel<NWB_BX_WARPS>;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, 32 * cw, smem, st>>>(bp, b->pc, (int)b->uni_A, (int)b->uni_B);
            CK(cudaGetLastError());
            b->launches += 1;
            return NWB_OK;
        }
        const size_t smem = NWB_BX_SMEM_PER_WARP(b->max_B) * (size_t)warps;
        CK(cudaFuncSetAttribute(nwb_batch_bx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nwb_batch_bx_kernel<<<grid, 32 * warps, smem, st>>>(bp, b->pc);
        CK(cudaGetLastError());
        b->launches += 1;
        return NWB_OK;
    }
    /* as many warps per SM as the per-warp shared memory (arrow ring + side string) allows, at most 12 */
    int warps = (int)((220 * 1024) / NWB_BATCH_SMEM_PER_WARP(b->max_B));
    if (warps > NWB_BATCH_WARPS) warps = NWB_BATCH_WARPS;
    if (warps < 1) return NWB_ERR_UNSUPPORTED;
    const long long nwarps = (long long)grid * warps;
    bp.bpitch = nwb_round_up((size_t)b->max_B + 1 + 64 + 256, 32);
    bp.scratch_per_warp = (b->max_strips > 1) ? (size_t)(b->max_strips - 1) * bp.bpitch : 0;
    if (bp.scratch_per_warp) {
        int rc = b->scratch.ensure((size_t)nwarps * bp.scratch_per_warp);
        if (rc != NWB_OK) return rc;
    }
    bp.tops = b->tops.p; bp.top_off = b->top_off.p + c0; bp.sides = b->sides.p; bp.side_off = b->side_off.p + c0;
    bp.n_pairs = c1 - c0; bp.m = b->m; bp.k = b->k; bp.d = b->d; bp.max_B = b->max_B;
    bp.arrows = b->arrows.p; bp.arrow_off = b->arrow_off.p + c0; bp.out_score = b->score.p + c0; bp.scratch = b->scratch.p;
    bp.out_branch = (b->flags & NWB_NO_BRANCH_COUNT) ? nullptr : b->branch.p + c0;
    const size_t smem = NWB_BATCH_SMEM_PER_WARP(b->max_B) * (size_t)warps;
    CK(cudaFuncSetAttribute(nwb_batch_pk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nwb_batch_pk_kernel<<<grid, 32 * warps, smem, st>>>(bp, b->pc);
    CK(cudaGetLastError());
    b->launches += 1;
    return NWB_OK;
}

extern "C" int nwb_batch_fetch(nwb_batch *b)
{
    if (!b || !b->ran) return NWB_ERR_INVALID;
    CK(cudaSetDevice(b->device));
    CK(cudaDeviceSynchronize());
    b->h_score.resize((size_t)b->n);
    b->h_branch.assign((size_t)b->n, 0u);
    if (b->n) {
        CK(cudaMemcpy(b->h_score.data(), b->score.p, (size_t)b->n * sizeof(int), cudaMemcpyDeviceToHost));
        if (!(b->flags & NWB_NO_BRANCH_COUNT))
            CK(cudaMemcpy(b->h_branch.data(), b->branch.p, (size_t)b->n * sizeof(unsigned), cudaMemcpyDeviceToHost));
        if (b->flags & NWB_WANT_COUNT) {
            b->h_count.resize((size_t)b->n);
            CK(cudaMemcpy(b->h_count.data(), b->count.p, (size_t)b->n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        }
        if (b->general && (b->flags & NWB_TRACK_ABS)) {
            b->h_abs.resize((size_t)b->n);
            CK(cudaMemcpy(b->h_abs.data(), b->gabs.p, (size_t)b->n * sizeof(int), cudaMemcpyDeviceToHost));
        }
        if (b->general && (b->flags & NWB_WANT_SCORES) && b->gscores_elems) {
            b->h_gscores.resize(b->gscores_elems);
            CK(cudaMemcpy(b->h_gscores.data(), b->gscores.p, b->gscores_elems * sizeof(int32_t), cudaMemcpyDeviceToHost));
        }
        if (b->flags & NWB_WANT_ARROWS_HOST) {
            b->h_arrows.resize(b->arrows_bytes);
            if (b->arrows_bytes) CK(cudaMemcpy(b->h_arrows.data(), b->arrows.p, b->arrows_bytes, cudaMemcpyDeviceToHost));
        }
    }
    b->fetched = true;
    return NWB_OK;
}

extern "C" int nwb_fill_batch(const char *tops, const int64_t *top_off, const char *sides, const int64_t *side_off,
                              int64_t n_pairs, int m, int k, int d, unsigned flags, int device, nwb_batch **out)
{
    if (!out) return NWB_ERR_INVALID;
    nwb_batch *b = nullptr;
    /* the strings go up chunk by chunk, overlapped with the kernels of the previous chunk */
    int rc = batch_create_impl(tops, top_off, sides, side_off, n_pairs, m, k, d, flags, device, false, &b);
    if (rc == NWB_OK) rc = nwb_batch_refill(b, tops, sides);
    if (rc == NWB_OK) rc = nwb_batch_fetch(b);
    if (rc != NWB_OK) { nwb_batch_free(b); *out = nullptr; return rc; }
    *out = b;
    return NWB_OK;
}

extern "C" int64_t nwb_batch_size(const nwb_batch *b) { return b ? b->n : 0; }
extern "C" int32_t nwb_batch_opt_score(const nwb_batch *b, int64_t pair)
{
    return (b && b->fetched && pair >= 0 && pair < b->n) ? b->h_score[(size_t)pair] : 0;
}
extern "C" uint32_t nwb_batch_branch_count(const nwb_batch *b, int64_t pair)
{
    return (b && b->fetched && pair >= 0 && pair < b->n) ? b->h_branch[(size_t)pair] : 0u;
}
extern "C" uint64_t nwb_batch_count_u64(const nwb_batch *b, int64_t pair)
{
    return (b && b->fetched && (b->flags & NWB_WANT_COUNT) && pair >= 0 && pair < b->n) ? b->h_count[(size_t)pair] : 0ull;
}
extern "C" const uint8_t *nwb_batch_arrow_rows(const nwb_batch *b, int64_t pair, size_t *pitch)
{
    if (!b || pair < 0 || pair >= b->n) return nullptr;
    const long long A = b->h_top_off[(size_t)pair + 1] - b->h_top_off[(size_t)pair];
    if (pitch) *pitch = (size_t)((A + 255) / 256 > 0 ? (A + 255) / 256 : 1) * 128;
    if (!b->fetched || b->h_arrows.empty()) return nullptr;
    return b->h_arrows.data() + b->h_arrow_off[(size_t)pair];
}
extern "C" int32_t nwb_batch_greatest_abs(const nwb_batch *b, int64_t pair)
{
    return (b && b->fetched && !b->h_abs.empty() && pair >= 0 && pair < b->n) ? b->h_abs[(size_t)pair] : 0;
}
extern "C" const int32_t *nwb_batch_score_rows(const nwb_batch *b, int64_t pair, size_t *pitch_elems)
{
    if (!b || pair < 0 || pair >= b->n) return nullptr;
    const long long A = b->h_top_off[(size_t)pair + 1] - b->h_top_off[(size_t)pair];
    if (pitch_elems) *pitch_elems = (size_t)((A + 255) / 256 > 0 ? (A + 255) / 256 : 1) * 256;
    if (!b->fetched || b->h_gscores.empty()) return nullptr;
    return b->h_gscores.data() + b->h_score_off[(size_t)pair];
}
extern "C" float nwb_batch_kernel_ms(const nwb_batch *b)
{
    if (!b || !b->ran || b->n == 0) return 0.f;
    cudaSetDevice(b->device);
    float ms = 0.f;
    if (cudaEventSynchronize(b->ev1) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, b->ev0, b->ev1) != cudaSuccess) return -1.f;
    return ms;
}
extern "C" const char *nwb_batch_kernel_name(const nwb_batch *b)
{
    if (!b) return "";
    if (b->general) return "nwb_batch_i32_kernel";
    if (b->use_bp) return "nwb_batch_bp_kernel";
    return b->use_cx ? "nwb_batch_cx_kernel" : (b->use_bx ? "nwb_batch_bx_kernel" : "nwb_batch_pk_kernel");
}
extern "C" int64_t nwb_batch_launches(const nwb_batch *b) { return b ? b->launches : 0; }
/* Digests of the whole batch, on the device (include/nwb.h): out[0..3] = sum over pairs p of
 * mix64(first_pair + p, x_p) with x_p = arrow digest / optimal score / branch count / alignment count. */
extern "C" int nwb_batch_digest(nwb_batch *b, int64_t first_pair, uint64_t out[4])
{
    if (!b || !out || !b->ran) return NWB_ERR_INVALID;
    out[0] = out[1] = out[2] = out[3] = 0;
    if (b->n == 0) return NWB_OK;
    CK(cudaSetDevice(b->device));
    CK(cudaDeviceSynchronize());
    int rc = b->digest.ensure(4);
    if (rc != NWB_OK) return rc;
    CK(cudaMemsetAsync(b->digest.p, 0, 4 * sizeof(unsigned long long), b->stream));
    nwb_batch_digest_kernel<<<b->sm_count * 8, 256, 0, b->stream>>>(
        b->arrows.p, b->arrow_off.p, b->top_off.p, b->side_off.p, b->n, first_pair, b->score.p,
        (b->flags & NWB_NO_BRANCH_COUNT) ? nullptr : b->branch.p, (b->flags & NWB_WANT_COUNT) ? b->count.p : nullptr,
        b->digest.p);
    CK(cudaGetLastError());
    b->launches += 1;
    unsigned long long h[4];
    CK(cudaMemcpyAsync(h, b->digest.p, sizeof(h), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    for (int i = 0; i < 4; i++) out[i] = h[i];
    return NWB_OK;
}
extern "C" void *nwb_batch_arrows_device(nwb_batch *b) { return b ? (void *)b->arrows.p : nullptr; }
