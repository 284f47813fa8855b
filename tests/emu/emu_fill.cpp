/*
 * emu_fill.cpp -- runs the REAL kernel sources (needleman-wunsch_b200/csrc/
 * *.cuh) on the CPU under the SIMT emulator.  TEST INFRASTRUCTURE ONLY: built
 * by tests/emu/Makefile into tests/emu/libnwb_emu.so, loaded only by tests.
 */
#define NWB_EMU 1
#include "emu_cuda.h"
#include "nwb_layout.h"
#include "nwb_fill_i32.cuh"
#include "nwb_fill_pk.cuh"
#include "nwb_fill_hx.cuh"
#include "nwb_fill_hz.cuh"
#include "nwb_fill_hy.cuh"
#include "nwb_count.cuh"
#include "nwb_batch.cuh"
#include "nwb_batch_bx.cuh"
#include "nwb_batch_bp.cuh"
#include "nwb_batch_count.cuh"
#include "nwb_batch_lcount.cuh"
#include "nwb_batch_i32.cuh"

#include <vector>

template <bool COUNT, bool SCORES, bool ABS, bool CNTMAT>
static void run_i32(unsigned grid, const NwbStripParams &p)
{
    emu_launch(grid, 32 * NWB_I32_WARPS, NWB_I32_SMEM_BYTES,
               [&]() { nwb_fill_i32_kernel<COUNT, SCORES, ABS, CNTMAT>(p); });
}

template <int K, int R, bool COUNT>
static void run_pk_emu(unsigned grid, int warps, const NwbStripParams &p, const NwbPkConsts &pc)
{
    emu_launch(grid, 32 * warps, (size_t)warps * NWB_PK_WARP_SMEM(K, R, COUNT),
               [&]() { nwb_fill_pk_kernel<K, R, COUNT>(p, pc); });
}

template <int M, int NW>
static bool emu_bp_go_w(int N, unsigned grid, int warps, const NwbBpParams &p)
{
    const size_t smem = NWB_BP_SMEM_BYTES(warps);
    switch (N) {
    case 0: emu_launch(grid, 32 * warps, smem, [&]() { nwb_batch_bp_kernel<M, 0, NW, false>(p); }); return true;
    case 1: if (M >= 1) { emu_launch(grid, 32 * warps, smem, [&]() { nwb_batch_bp_kernel<M, (M >= 1 ? 1 : 0), NW, false>(p); }); return true; } break;
    case 2: if (M >= 2) { emu_launch(grid, 32 * warps, smem, [&]() { nwb_batch_bp_kernel<M, (M >= 2 ? 2 : 0), NW, false>(p); }); return true; } break;
    case 3: if (M >= 3) { emu_launch(grid, 32 * warps, smem, [&]() { nwb_batch_bp_kernel<M, (M >= 3 ? 3 : 0), NW, false>(p); }); return true; } break;
    }
    return false;
}
template <int M>
static bool emu_bp_go(int N, int nw, unsigned grid, int warps, const NwbBpParams &p)
{
    if (nw == 2) return emu_bp_go_w<M, 2>(N, grid, warps, p);
    if (nw == 4) return emu_bp_go_w<M, 4>(N, grid, warps, p);
    return emu_bp_go_w<M, 8>(N, grid, warps, p);
}

extern "C" {

struct emu_out {
    int opt_score;
    unsigned branch_count;
    int greatest_abs;
    int pad;
    unsigned long long count;
    unsigned long long pitch;
    unsigned long long spitch;
    unsigned long long dig_row, dig_col; /* dense count sweep (cpl 8): NwbDevSummary.dig_row / dig_col */
};

/* flags: 1 scores, 2 count, 8 abs, 0x20 cntmat (same bits as include/nwb.h).
 * arrows: B*pitch bytes (pitch from emu_pitch_i32), scores/cntmat: B*spitch. */
size_t emu_pitch_i32(int A, int B) { return nwb_make_layout(A, B, NWB_KIND_I32, 0, NWB_I32_STRIP_W).pitch; }
size_t emu_spitch_i32(int A, int B) { return nwb_make_layout(A, B, NWB_KIND_I32, 0, NWB_I32_STRIP_W).spitch; }


int emu_fill_i32(const char *top, int A, const char *side, int B, int m, int k, int d,
                 unsigned flags, unsigned grid, int split, /* emulate a 2-GPU strip split at this strip (0 = none) */
                 uint8_t *arrows, int32_t *scores, unsigned long long *cntmat, emu_out *out)
{
    NwbLayout L = nwb_make_layout(A, B, NWB_KIND_I32, 0, NWB_I32_STRIP_W);
    std::vector<int32_t> bnd_s((size_t)L.n_strips * L.bpitch, 0x7f7f7f7f);
    std::vector<unsigned long long> bnd_c((size_t)L.n_strips * L.bpitch, 0xdeadbeefULL);
    std::vector<int> progress((size_t)L.n_strips, 0);
    NwbDevSummary sum;
    memset(&sum, 0, sizeof(sum));

    NwbStripParams p;
    memset(&p, 0, sizeof(p));
    p.top = (const uint8_t *)top;
    p.side = (const uint8_t *)side;
    p.A = A; p.B = B; p.m = m; p.k = k; p.d = d;
    p.n_strips = L.n_strips;
    p.strip_begin = 0;
    p.strip_end = L.n_strips;
    p.arrows = arrows;
    p.pitch = L.pitch;
    p.scores = scores;
    p.cntmat = cntmat;
    p.spitch = L.spitch;
    p.bnd_s = bnd_s.data();
    p.bnd_c = bnd_c.data();
    p.bpitch = L.bpitch;
    p.progress = progress.data();
    p.summary = &sum;

    const bool C = flags & 2, S = flags & 1, AB = flags & 8, CM = flags & 0x20;
    auto launch = [&](const NwbStripParams &q) {
        if (C) {
            if (S) { if (AB) { if (CM) run_i32<true, true, true, true>(grid, q); else run_i32<true, true, true, false>(grid, q); }
                     else { if (CM) run_i32<true, true, false, true>(grid, q); else run_i32<true, true, false, false>(grid, q); } }
            else { if (AB) { if (CM) run_i32<true, false, true, true>(grid, q); else run_i32<true, false, true, false>(grid, q); }
                   else { if (CM) run_i32<true, false, false, true>(grid, q); else run_i32<true, false, false, false>(grid, q); } }
        } else {
            if (S) { if (AB) run_i32<false, true, true, false>(grid, q); else run_i32<false, true, false, false>(grid, q); }
            else { if (AB) run_i32<false, false, true, false>(grid, q); else run_i32<false, false, false, false>(grid, q); }
        }
    };
    if (split > 0 && split < L.n_strips) {
        /* two "GPUs" run one after the other: the first publishes its last
         * strip's boundary through the out_* pointers into the second's arrays */
        NwbStripParams p0 = p, p1 = p;
        std::vector<int32_t> inbox_s(L.bpitch, 0x7f7f7f7f);
        std::vector<unsigned long long> inbox_c(L.bpitch, 0xdeadbeefULL);
        int inbox_flag = 0;
        p0.strip_end = split;
        p0.out_bnd_s = inbox_s.data();
        p0.out_bnd_c = inbox_c.data();
        p0.out_progress = &inbox_flag;
        /* the second "GPU" has its own (zeroed) progress words and boundary arrays, as every nwb_plan has */
        std::vector<int32_t> bnd1_s((size_t)L.n_strips * L.bpitch, 0x7f7f7f7f);
        std::vector<unsigned long long> bnd1_c((size_t)L.n_strips * L.bpitch, 0xdeadbeefULL);
        std::vector<int> progress1((size_t)L.n_strips, 0);
        p1.strip_begin = split;
        p1.in_bnd_s = inbox_s.data();
        p1.in_bnd_c = inbox_c.data();
        p1.in_progress = &inbox_flag;
        p1.bnd_s = bnd1_s.data();
        p1.bnd_c = bnd1_c.data();
        p1.progress = progress1.data();
        launch(p0);
        launch(p1);
    } else {
        launch(p);
    }
    out->opt_score = sum.opt_score;
    out->branch_count = sum.branch_count;
    out->greatest_abs = sum.greatest_abs;
    out->count = sum.count;
    out->pitch = L.pitch;
    out->spitch = L.spitch;
    return 0;
}

size_t emu_pitch_pk(int A, int B, int K) { return nwb_make_layout(A, B, NWB_KIND_PK, K, 64 * K).pitch; }
int emu_pk_supported(int m, int k, int d) { return nwb_pk_supported(m, k, d, nullptr) ? 1 : 0; }

int emu_hx_supported(int m, int k, int d)
{
    NwbPkConsts pc;
    return (nwb_pk_supported(m, k, d, &pc) && nwb_hx_supported(pc)) ? 1 : 0;
}

/* hx != 0: the sweeping + flush warp variant (nwb_fill_hx.cuh; K = 4, R = 2, no fused count).
 * count: 0 none, 1 fused into nwb_fill_pk_kernel, 2 / 4 / 8 second sweep over the arrow codes with that
 * many cells per lane (nwb_count.cuh). */
int emu_fill_pk(const char *top, int A, const char *side, int B, int m, int k, int d, int K, int R, int count,
                unsigned grid, int warps, int split, int hx, uint8_t *arrows, emu_out *out)
{
    NwbPkConsts pc;
    if (!nwb_pk_supported(m, k, d, &pc)) return -5;
    if (hx && (!nwb_hx_supported(pc) || K != 4 || R != 2 || count == 1)) return -6;
    if (count == 1 && K != 4) return -6; /* the fused count exists for K = 4 only (nwb_pk_launch rejects it too) */
    if (count >= 2 && (K != 4 || (count != 2 && count != 4 && count != 8))) return -6;
    NwbLayout L = nwb_make_layout(A, B, NWB_KIND_PK, K, 64 * K);
    std::vector<uint32_t> bnd_w((size_t)L.n_strips * L.bpitch, 0u);
    std::vector<int> progress((size_t)L.n_strips, 0);
    NwbDevSummary sum;
    memset(&sum, 0, sizeof(sum));
    NwbStripParams p;
    memset(&p, 0, sizeof(p));
    p.top = (const uint8_t *)top;
    p.side = (const uint8_t *)side;
    p.A = A; p.B = B; p.m = m; p.k = k; p.d = d;
    p.n_strips = L.n_strips;
    p.strip_begin = 0;
    p.strip_end = L.n_strips;
    p.arrows = arrows;
    p.pitch = L.pitch;
    p.bnd_w = bnd_w.data();
    std::vector<unsigned long long> bnd_cc((size_t)L.n_strips * 2 * L.bpitch, 0ull);
    p.bnd_c = bnd_cc.data();
    p.bpitch = L.bpitch;
    p.progress = progress.data();
    p.summary = &sum;
    p.count_branches = 1;
    std::vector<uint16_t> side_pre(NWB_PK_SPRE_LEN(B), 0x1234);
    emu_launch(2, 64, 0, [&]() { nwb_pk_prep_side_kernel((const uint8_t *)side, B, pc.shift, side_pre.data()); });
    p.side_pre = side_pre.data();
    const uint32_t *last_stream = nullptr; /* the stream consumed by the last strip, when it is not in bnd_w */
    if (hx >= 4) p.hx_spb = hx - 3; /* hx = 4, 5, 6: queue mode of nwb_fill_hx_kernel with 1, 2, 3 adjacent strips per block */
    auto launch = [&](const NwbStripParams &q) {
        sum.ticket = 0; /* the hx blocks draw their ids from here; nwb_plan_run() zeroes the summary block per launch */
        if (hx == 3) { /* sweeping + packing + flush warps, two strips per block (nwb_fill_hz.cuh) */
            if (q.publish_rows) emu_launch(grid, 32 * NWB_HZ_WARPS, NWB_HZ_SMEM_BYTES, [&]() { nwb_fill_hz_kernel<true>(q, pc); });
            else emu_launch(grid, 32 * NWB_HZ_WARPS, NWB_HZ_SMEM_BYTES, [&]() { nwb_fill_hz_kernel<false>(q, pc); });
        } else if (hx == 2) { /* one row of skew per virtual lane (nwb_fill_hy.cuh) */
            if (q.publish_rows) emu_launch(grid, 32 * NWB_HX_WARPS, NWB_HX_SMEM_BYTES, [&]() { nwb_fill_hy_kernel<true>(q, pc); });
            else emu_launch(grid, 32 * NWB_HX_WARPS, NWB_HX_SMEM_BYTES, [&]() { nwb_fill_hy_kernel<false>(q, pc); });
        } else if (hx >= 4) {
            emu_launch(grid, 32 * NWB_HX_WARPS, NWB_HX_SMEM_BYTES, [&]() { nwb_fill_hx_kernel<false, true>(q, pc); });
        } else if (hx) {
            if (q.publish_rows) emu_launch(grid, 32 * NWB_HX_WARPS, NWB_HX_SMEM_BYTES, [&]() { nwb_fill_hx_kernel<true>(q, pc); });
            else emu_launch(grid, 32 * NWB_HX_WARPS, NWB_HX_SMEM_BYTES, [&]() { nwb_fill_hx_kernel<false>(q, pc); });
        } else if (count == 1) {
            if (R == 2) run_pk_emu<4, 2, true>(grid, warps, q, pc);
            else run_pk_emu<4, 1, true>(grid, warps, q, pc);
        } else if (R == 2) {
            if (K == 1) run_pk_emu<1, 2, false>(grid, warps, q, pc);
            else if (K == 2) run_pk_emu<2, 2, false>(grid, warps, q, pc);
            else run_pk_emu<4, 2, false>(grid, warps, q, pc);
        } else {
            if (K == 1) run_pk_emu<1, 1, false>(grid, warps, q, pc);
            else if (K == 2) run_pk_emu<2, 1, false>(grid, warps, q, pc);
            else run_pk_emu<4, 1, false>(grid, warps, q, pc);
        }
    };
    /* count >= 2: second sweep over the arrow codes with count cells per lane (8, 4 or 2); split_strip in
     * 256-column fill strips, two launches chained through the inbox like two GPUs */
    std::vector<unsigned long long> cnt_bnd;
    /* cpl == 8 after the hx kernel: the sweep also waits on the rows the flush warps published (it trails the
     * fill on a second stream on the GPU; here the fill has finished, so every wait must already be satisfied) */
    const int *prog0 = nullptr, *prog1 = nullptr;
    auto run_count = [&](int cpl, int split_strip, unsigned long long *outbox, const unsigned long long *inbox) {
        const int wc = 32 * cpl, ratio = 256 / wc;
        NwbCountParams cp;
        memset(&cp, 0, sizeof(cp));
        cp.arrows = arrows; cp.pitch = L.pitch; cp.A = A; cp.B = B;
        cp.n_strips = (A + wc - 1) / wc;
        cp.bpitch = L.bpitch;
        cp.summary = &sum;
        cnt_bnd.assign((size_t)cp.n_strips * 2 * L.bpitch, 0ull);
        auto go = [&](const NwbCountParams &q) {
            if (cpl == 2) emu_launch(grid, 32 * NWB_CNT_WARPS, NWB_CNT_SMEM_BYTES, [&]() { nwb_count_kernel<2, false>(q); });
            else if (cpl == 4) emu_launch(grid, 32 * NWB_CNT_WARPS, NWB_CNT_SMEM_BYTES, [&]() { nwb_count_kernel<4, false>(q); });
            else emu_launch(grid, 32 * NWB_CNT_WARPS, NWB_CNT_SMEM_BYTES, [&]() { nwb_count_kernel<8, true>(q); }); /* with digests */
        };
        if (split_strip > 0) {
            int sb = split_strip * ratio;
            if (sb > cp.n_strips) sb = cp.n_strips;
            NwbCountParams c0 = cp, c1 = cp;
            c0.strip_begin = 0; c0.strip_end = sb; c0.bnd_c = cnt_bnd.data(); c0.out_bnd_c = outbox;
            c1.strip_begin = sb; c1.strip_end = cp.n_strips; c1.bnd_c = cnt_bnd.data() + (size_t)sb * 2 * L.bpitch;
            c1.in_bnd_c = inbox;
            if (cpl == 8) { c0.fill_progress = prog0; c1.fill_progress = prog1; }
            if (c0.strip_end > c0.strip_begin) go(c0);
            if (c1.strip_end > c1.strip_begin) go(c1);
        } else {
            cp.strip_begin = 0; cp.strip_end = cp.n_strips; cp.bnd_c = cnt_bnd.data();
            if (cpl == 8) cp.fill_progress = prog0;
            go(cp);
        }
    };
    if (split > 0 && split < L.n_strips) {
        NwbStripParams p0 = p, p1 = p;
        std::vector<uint32_t> inbox_w(L.bpitch, 0u);
        std::vector<unsigned long long> inbox_cc(2 * L.bpitch, 0ull), bnd1c((size_t)L.n_strips * 2 * L.bpitch, 0ull);
        int inbox_flag = 0;
        std::vector<uint32_t> bnd1((size_t)L.n_strips * L.bpitch, 0u);
        std::vector<int> prog1v((size_t)L.n_strips, 0);
        if (hx && hx < 4 && count >= 2) { p0.publish_rows = 1; p1.publish_rows = 1; prog0 = progress.data(); prog1 = prog1v.data(); }
        p0.strip_end = split;
        p0.out_bnd_w = inbox_w.data();
        p0.out_bnd_c = inbox_cc.data();
        p0.out_progress = &inbox_flag;
        p1.strip_begin = split;
        p1.in_bnd_w = inbox_w.data();
        p1.in_bnd_c = inbox_cc.data();
        p1.bnd_c = bnd1c.data();
        p1.in_progress = &inbox_flag;
        p1.bnd_w = bnd1.data();
        p1.progress = prog1v.data();
        /* queue mode in a pipelined strip group: the last local strip looks at the right neighbour's acknowledgement
         * word before its first store into the neighbour's inbox (here: already granted) */
        uint32_t ack_word = 9u;
        if (hx >= 4) { p0.gate_ack = &ack_word; p0.gate_need = 7u; }
        launch(p0);
        launch(p1);
        if (count >= 2) run_count(count, split, p0.out_bnd_c, p1.in_bnd_c);
        last_stream = (split == L.n_strips - 1) ? inbox_w.data() : bnd1.data() + (size_t)(L.n_strips - 2 - split) * L.bpitch;
        emu_launch(2, 64, 0, [&]() { nwb_pk_stream_sum_kernel(last_stream, B, R, &sum.rsum); });
    } else {
        if (hx && hx < 4 && count >= 2) { p.publish_rows = 1; prog0 = progress.data(); }
        launch(p);
        if (count >= 2) run_count(count, 0, nullptr, nullptr);
        if (L.n_strips >= 2) {
            last_stream = bnd_w.data() + (size_t)(L.n_strips - 2) * L.bpitch;
            emu_launch(2, 64, 0, [&]() { nwb_pk_stream_sum_kernel(last_stream, B, R, &sum.rsum); });
        }
    }
    /* the fused counter (flush) and the stand-alone pass over the finished table must agree */
    unsigned branches = 0;
    emu_launch(3, 64, 0, [&]() { nwb_branch_count_kernel(arrows, L.pitch, A, B, 0, A, &branches); });
    if (branches != sum.branch_count) return -77;
    out->opt_score = (int)(sum.rsum - (long long)d * ((long long)A + B));
    out->branch_count = sum.branch_count;
    out->greatest_abs = 0;
    out->count = sum.count;
    out->pitch = L.pitch;
    out->spitch = 0;
    out->dig_row = sum.dig_row;
    out->dig_col = sum.dig_col;
    return 0;
}

/* The sparse backward count (nwb_count_sparse.cuh) over a finished nibble table in the include/nwb.h layout.
 * res = {count, state (NWB_SPC_*), rows visited}. */
int emu_sparse_count(const uint8_t *arrows, size_t pitch, int A, int B, int mode, int min_col, unsigned long long *res)
{
    unsigned long long count = 0ull;
    int state = 0;
    unsigned rows = 0;
    NwbSparseCountParams sc;
    memset(&sc, 0, sizeof(sc));
    sc.arrows = arrows; sc.pitch = pitch; sc.A = A; sc.B = B;
    sc.out_count = &count; sc.out_state = &state; sc.out_rows = &rows;
    sc.mode = mode;
    sc.min_col = min_col;
    emu_launch(1, 32, 0, [&]() { nwb_sparse_count_kernel(sc); });
    res[0] = count; res[1] = (unsigned long long)state; res[2] = rows;
    return 0;
}

/* nwb_arrow_digest_kernel over words [w_begin, w_end) of a nibble table */
unsigned long long emu_arrow_digest(const uint8_t *arrows, size_t pitch, int A, int B, int w_begin, int w_end, unsigned grid)
{
    unsigned long long out = 0ull;
    emu_launch(grid, 256, 0, [&]() { nwb_arrow_digest_kernel(arrows, pitch, A, B, w_begin, w_end, &out); });
    return out;
}

/* ONE rank of a column-strip group under the emulator (tests/test_dist_gloo.py: one process per rank, the
 * boundary stream travels between the processes the way it crosses NVLink between GPUs).  The rank sweeps
 * strips [begin, end) given by nwb_rank_strip_range(); `inbox` = the bpitch words its first strip consumes
 * (written by rank - 1, NULL for rank 0); `outbox` = the words its last strip publishes for rank + 1 (NULL
 * for the last rank).  hx != 0: nwb_fill_hx_kernel, else nwb_fill_pk_kernel<4, 2>.  info = {strip_begin,
 * strip_end, n_strips, bpitch}; partial_r = this rank's share of sum_i u(i,B) as in nwb_plan_summary(). */
size_t emu_bpitch_pk(int A, int B) { return nwb_make_layout(A, B, NWB_KIND_PK, 4, 256).bpitch; }

int emu_fill_pk_rank(const char *top, int A, const char *side, int B, int m, int k, int d, unsigned grid, int hx,
                     int rank, int world, uint32_t *inbox, uint32_t *outbox, uint8_t *arrows,
                     long long *partial_r, unsigned *branches, int *info)
{
    const int K = 4, R = 2;
    NwbPkConsts pc;
    if (!nwb_pk_supported(m, k, d, &pc)) return -5;
    if (hx && !nwb_hx_supported(pc)) return -6;
    NwbLayout L = nwb_make_layout(A, B, NWB_KIND_PK, K, 64 * K);
    int sb, se;
    nwb_rank_strip_range(L.n_strips, rank, world, &sb, &se);
    info[0] = sb; info[1] = se; info[2] = L.n_strips; info[3] = (int)L.bpitch;
    *partial_r = 0;
    *branches = 0;
    if (se <= sb) return 0;
    if ((sb > 0 && !inbox) || (se < L.n_strips && !outbox)) return -1;
    std::vector<uint32_t> bnd_w((size_t)(se - sb) * L.bpitch, 0u);
    std::vector<int> progress((size_t)(se - sb), 0);
    NwbDevSummary sum;
    memset(&sum, 0, sizeof(sum));
    NwbStripParams p;
    memset(&p, 0, sizeof(p));
    p.top = (const uint8_t *)top;
    p.side = (const uint8_t *)side;
    p.A = A; p.B = B; p.m = m; p.k = k; p.d = d;
    p.n_strips = L.n_strips;
    p.strip_begin = sb;
    p.strip_end = se;
    p.arrows = arrows;
    p.pitch = L.pitch;
    p.bnd_w = bnd_w.data();
    p.bpitch = L.bpitch;
    p.progress = progress.data();
    p.summary = &sum;
    p.count_branches = 1;
    p.in_bnd_w = inbox;
    p.out_bnd_w = outbox;
    std::vector<uint16_t> side_pre(NWB_PK_SPRE_LEN(B), 0x1234);
    emu_launch(2, 64, 0, [&]() { nwb_pk_prep_side_kernel((const uint8_t *)side, B, pc.shift, side_pre.data()); });
    p.side_pre = side_pre.data();
    if (hx == 3) emu_launch(grid, 32 * NWB_HZ_WARPS, NWB_HZ_SMEM_BYTES, [&]() { nwb_fill_hz_kernel<false>(p, pc); });
    else if (hx == 2) emu_launch(grid, 32 * NWB_HX_WARPS, NWB_HX_SMEM_BYTES, [&]() { nwb_fill_hy_kernel<false>(p, pc); });
    else if (hx >= 4) { /* queue mode: ticketed blocks of hx - 3 adjacent strips; a rank that publishes checks its neighbour's
                         * acknowledgement word first (the caller has waited for the acknowledgement: granted) */
        uint32_t ack_word = 3u;
        p.hx_spb = hx - 3;
        if (se < L.n_strips) { p.gate_ack = &ack_word; p.gate_need = 3u; }
        emu_launch(grid, 32 * NWB_HX_WARPS, NWB_HX_SMEM_BYTES, [&]() { nwb_fill_hx_kernel<false, true>(p, pc); });
    }
    else if (hx) emu_launch(grid, 32 * NWB_HX_WARPS, NWB_HX_SMEM_BYTES, [&]() { nwb_fill_hx_kernel<false>(p, pc); });
    else run_pk_emu<4, 2, false>(grid, 4, p, pc);
    { /* the counter fused into the flush and the stand-alone pass over this rank's columns must agree */
        unsigned bc = 0;
        const int c0 = sb * 64 * K, c1 = (se * 64 * K < A) ? se * 64 * K : A;
        emu_launch(3, 64, 0, [&]() { nwb_branch_count_kernel(arrows, L.pitch, A, B, c0, c1, &bc); });
        if (bc != sum.branch_count) return -77;
    }
    if (L.n_strips >= 2 && se == L.n_strips) {
        const uint32_t *stream = (sb == L.n_strips - 1) ? inbox : bnd_w.data() + (size_t)(L.n_strips - 2 - sb) * L.bpitch;
        emu_launch(2, 64, 0, [&]() { nwb_pk_stream_sum_kernel(stream, B, R, &sum.rsum); });
    }
    *partial_r = sum.rsum;
    *branches = sum.branch_count;
    return 0;
}

/* nwb_batch_lcount_kernel (one thread per pair, sparse backward sweep) followed, as in batch_count_pass(), by the
 * dense nwb_batch_count_kernel over the pairs it gave up on.  *n_fallback reports how many those were. */
int emu_batch_lcount(const long long *top_off, const long long *side_off, long long n, unsigned grid, int warps,
                     const uint8_t *arrows, const long long *arrow_off, unsigned long long *counts, long long *n_fallback)
{
    std::vector<long long> fb((size_t)n + 1, -1);
    unsigned fbn = 0;
    NwbLaneCountParams lp;
    memset(&lp, 0, sizeof(lp));
    lp.top_off = top_off; lp.side_off = side_off; lp.n_pairs = n; lp.arrows = arrows; lp.arrow_off = arrow_off;
    lp.out_count = counts; lp.fb_list = fb.data(); lp.fb_count = &fbn;
    emu_launch(grid, 32 * warps, 0, [&]() { nwb_batch_lcount_kernel(lp); });
    if (n_fallback) *n_fallback = fbn;
    if (fbn) {
        NwbBatchCountParams cp;
        memset(&cp, 0, sizeof(cp));
        int maxB = 0, maxS = 1;
        for (long long p = 0; p < n; p++) {
            const long long A = top_off[p + 1] - top_off[p], B = side_off[p + 1] - side_off[p];
            if ((A + 255) / 256 > maxS) maxS = (int)((A + 255) / 256);
            if (B > maxB) maxB = (int)B;
        }
        cp.top_off = top_off; cp.side_off = side_off; cp.n_pairs = n;
        cp.arrows = arrows; cp.arrow_off = arrow_off; cp.out_count = counts;
        std::vector<unsigned long long> scratch;
        if (maxS > 1) {
            cp.scratch_per_warp = nwb_round_up((size_t)maxB + 1, 16);
            scratch.assign((size_t)grid * NWB_BCNT_WARPS * cp.scratch_per_warp, 0xdeadbeefdeadbeefULL);
            cp.scratch = scratch.data();
        }
        cp.pair_list = fb.data(); cp.pair_count = &fbn;
        emu_launch(grid, 32 * NWB_BCNT_WARPS, (size_t)NWB_BCNT_SMEM_PER_WARP * NWB_BCNT_WARPS, [&]() { nwb_batch_count_kernel(cp); });
    }
    return 0;
}

/* the batch count pass (nwb_batch_count.cuh) over tables a batch fill has written */
int emu_batch_count(const long long *top_off, const long long *side_off, long long n, unsigned grid,
                    const uint8_t *arrows, const long long *arrow_off, unsigned long long *counts)
{
    NwbBatchCountParams cp;
    memset(&cp, 0, sizeof(cp));
    int maxB = 0, maxS = 1;
    for (long long p = 0; p < n; p++) {
        const long long A = top_off[p + 1] - top_off[p], B = side_off[p + 1] - side_off[p];
        if ((A + 255) / 256 > maxS) maxS = (int)((A + 255) / 256);
        if (B > maxB) maxB = (int)B;
    }
    cp.top_off = top_off; cp.side_off = side_off; cp.n_pairs = n;
    cp.arrows = arrows; cp.arrow_off = arrow_off; cp.out_count = counts;
    std::vector<unsigned long long> scratch;
    if (maxS > 1) {
        cp.scratch_per_warp = nwb_round_up((size_t)maxB + 1, 16);
        scratch.assign((size_t)grid * NWB_BCNT_WARPS * cp.scratch_per_warp, 0xdeadbeefdeadbeefULL);
        cp.scratch = scratch.data();
    }
    /* as batch_count_pass() in csrc/nwb_batch_api.inl: uniform one-strip batches are swept back to back */
    bool uniform = n > 0;
    const long long A0 = n > 0 ? top_off[1] - top_off[0] : 0, B0 = n > 0 ? side_off[1] - side_off[0] : 0;
    for (long long p = 1; p < n; p++)
        if (top_off[p + 1] - top_off[p] != A0 || side_off[p + 1] - side_off[p] != B0) uniform = false;
    if (nwb_bcount_chain_usable(uniform, A0, B0, n, (long long)grid * NWB_BCNT_WARPS))
        emu_launch(grid, 32 * NWB_BCNT_WARPS, (size_t)NWB_BCNT_SMEM_PER_WARP * NWB_BCNT_WARPS,
                   [&]() { nwb_batch_count_chain_kernel(cp, (int)A0, (int)B0); });
    else
        emu_launch(grid, 32 * NWB_BCNT_WARPS, (size_t)NWB_BCNT_SMEM_PER_WARP * NWB_BCNT_WARPS, [&]() { nwb_batch_count_kernel(cp); });
    return 0;
}

/* batch kernel under the emulator: arrows = concatenated per-pair tables (offsets returned in arrow_off) */
/* bx: 0 = nwb_batch_pk_kernel (one pair per warp), 1 = nwb_batch_bx_kernel (two pairs per warp; -6 when the
 * batch does not qualify), 2 = nwb_batch_cx_kernel (uniform shapes, pairs back to back; -6 likewise), -1 =
 * whatever nwb_batch_run() would pick.  *used_bx reports the choice (0, 1 or 2). */
/* nwb_batch_i32_kernel<scores, abs>: any m / k / d.  out_abs / scores may be NULL. */
int emu_fill_batch_i32(const char *tops, const long long *top_off, const char *sides, const long long *side_off,
                       long long n, int m, int k, int d, unsigned grid, uint8_t *arrows, long long *arrow_off,
                       int *out_score, unsigned *out_branch, int *out_abs, int32_t *scores, long long *score_off)
{
    int maxB = 0, maxS = 1;
    long long aoff = 0, soff = 0;
    for (long long p = 0; p < n; p++) {
        const long long A = top_off[p + 1] - top_off[p], B = side_off[p + 1] - side_off[p];
        const int ns = (int)((A + 255) / 256);
        if (ns > maxS) maxS = ns;
        if (B > maxB) maxB = (int)B;
        arrow_off[p] = aoff;
        aoff += (long long)(ns > 0 ? ns : 1) * 128 * B;
        if (score_off) { score_off[p] = soff; soff += (long long)(ns > 0 ? ns : 1) * 256 * B; }
    }
    arrow_off[n] = aoff;
    if (score_off) score_off[n] = soff;
    const size_t nwarps = (size_t)grid * NWB_BI32_WARPS;
    NwbBatchI32Params gp;
    memset(&gp, 0, sizeof(gp));
    gp.tops = (const uint8_t *)tops; gp.top_off = top_off; gp.sides = (const uint8_t *)sides; gp.side_off = side_off;
    gp.n_pairs = n; gp.m = m; gp.k = k; gp.d = d;
    gp.arrows = arrows; gp.arrow_off = arrow_off; gp.scores = scores; gp.score_off = score_off;
    gp.out_score = out_score; gp.out_branch = out_branch; gp.out_abs = out_abs;
    gp.bpitch = nwb_round_up((size_t)maxB + 2, 32);
    std::vector<int32_t> bnd(nwarps * (size_t)maxS * gp.bpitch, 0x7f7f7f7f);
    std::vector<int> prog(nwarps * (size_t)maxS, 12345);
    std::vector<NwbDevSummary> wsum(nwarps);
    gp.bnd_s = bnd.data(); gp.progress = prog.data(); gp.max_strips = maxS; gp.wsum = wsum.data();
    const size_t smem = (size_t)NWB_BI32_WARPS * NWB_I32_STAGE_WORDS * 4;
    if (scores) {
        if (out_abs) emu_launch(grid, 32 * NWB_BI32_WARPS, smem, [&]() { nwb_batch_i32_kernel<true, true>(gp); });
        else emu_launch(grid, 32 * NWB_BI32_WARPS, smem, [&]() { nwb_batch_i32_kernel<true, false>(gp); });
    } else {
        if (out_abs) emu_launch(grid, 32 * NWB_BI32_WARPS, smem, [&]() { nwb_batch_i32_kernel<false, true>(gp); });
        else emu_launch(grid, 32 * NWB_BI32_WARPS, smem, [&]() { nwb_batch_i32_kernel<false, false>(gp); });
    }
    return 0;
}

int emu_fill_batch(const char *tops, const long long *top_off, const char *sides, const long long *side_off,
                   long long n, int m, int k, int d, unsigned grid, int bx, uint8_t *arrows, long long *arrow_off,
                   int *scores, unsigned *branches, int *used_bx)
{
    NwbPkConsts pc;
    if (!nwb_pk_supported(m, k, d, &pc)) return -5;
    NwbBatchParams bp;
    memset(&bp, 0, sizeof(bp));
    int maxB = 0, maxS = 1;
    long long aoff = 0;
    for (long long p = 0; p < n; p++) {
        const long long A = top_off[p + 1] - top_off[p], B = side_off[p + 1] - side_off[p];
        const int ns = (int)((A + 255) / 256);
        if (ns > maxS) maxS = ns;
        if (B > maxB) maxB = (int)B;
        arrow_off[p] = aoff;
        aoff += (long long)(ns > 0 ? ns : 1) * 128 * B;
    }
    arrow_off[n] = aoff;
    long long maxA = 0;
    for (long long p = 0; p < n; p++)
        if (top_off[p + 1] - top_off[p] > maxA) maxA = top_off[p + 1] - top_off[p];
    const bool can_bx = nwb_bx_usable(pc, maxA, maxB);
    bool uniform = n > 0;
    for (long long p = 1; p < n; p++)
        if (top_off[p + 1] - top_off[p] != top_off[1] - top_off[0] || side_off[p + 1] - side_off[p] != side_off[1] - side_off[0])
            uniform = false;
    const bool can_cx = can_bx && n > 0 && nwb_cx_usable(pc, uniform, top_off[1] - top_off[0], (int)(side_off[1] - side_off[0]));
    if ((bx == 1 && !can_bx) || (bx == 2 && !can_cx)) return -6;
    const bool use_cx = (bx == 2 || bx == -1) && can_cx;
    const bool use_bx = (bx != 0) && can_bx;
    if (used_bx) *used_bx = use_cx ? 2 : (use_bx ? 1 : 0);
    if (use_bx) {
        bp.tops = (const uint8_t *)tops; bp.top_off = top_off; bp.sides = (const uint8_t *)sides; bp.side_off = side_off;
        bp.n_pairs = n; bp.m = m; bp.k = k; bp.d = d; bp.max_B = maxB;
        bp.arrows = arrows; bp.arrow_off = arrow_off; bp.out_score = scores; bp.out_branch = branches;
        if (use_cx) {
            const int A = (int)(top_off[1] - top_off[0]), B = (int)(side_off[1] - side_off[0]);
            if (getenv("NWB_CX_WARPS") && atoi(getenv("NWB_CX_WARPS")) == 16)
                emu_launch(grid, 32 * 16, NWB_CX_SMEM_PER_WARP(B, 16) * 16, [&]() { nwb_batch_cx_kernel<16>(bp, pc, A, B); });
            else
                emu_launch(grid, 32 * NWB_BX_WARPS, NWB_CX_SMEM_PER_WARP(B, NWB_BX_WARPS) * NWB_BX_WARPS,
                           [&]() { nwb_batch_cx_kernel<NWB_BX_WARPS>(bp, pc, A, B); });
        } else {
            emu_launch(grid, 32 * NWB_BX_WARPS, NWB_BX_SMEM_PER_WARP(maxB) * NWB_BX_WARPS, [&]() { nwb_batch_bx_kernel(bp, pc); });
        }
        return 0;
    }
    const long long nwarps = (long long)grid * NWB_BATCH_WARPS;
    bp.bpitch = nwb_round_up((size_t)maxB + 1 + 64 + 256, 32);
    bp.scratch_per_warp = (maxS > 1) ? (size_t)(maxS - 1) * bp.bpitch : 0;
    std::vector<uint32_t> scratch((size_t)nwarps * bp.scratch_per_warp + 1, 0xdeadbeefu);
    bp.tops = (const uint8_t *)tops; bp.top_off = top_off; bp.sides = (const uint8_t *)sides; bp.side_off = side_off;
    bp.n_pairs = n; bp.m = m; bp.k = k; bp.d = d; bp.max_B = maxB;
    bp.arrows = arrows; bp.arrow_off = arrow_off; bp.out_score = scores; bp.scratch = scratch.data();
    bp.out_branch = branches;
    emu_launch(grid, 32 * NWB_BATCH_WARPS, NWB_BATCH_SMEM_PER_WARP(maxB) * NWB_BATCH_WARPS,
               [&]() { nwb_batch_pk_kernel(bp, pc); });
    return 0;
}

/* nwb_batch_bp_kernel<M, N> (bit-parallel rows, one thread per pair) followed, as in batch_fill_pass(), by
 * nwb_batch_pk_kernel over the pairs it left on its list (more than four distinct letters in the top string).
 * *n_fallback reports how many those were.  -6: the batch does not qualify. */
int emu_fill_batch_bp(const char *tops, const long long *top_off, const char *sides, const long long *side_off,
                      long long n, int m, int k, int d, unsigned grid, int warps, int force_words, uint8_t *arrows,
                      long long *arrow_off, int *scores, unsigned *branches, long long *n_fallback)
{
    NwbPkConsts pc;
    if (!nwb_pk_supported(m, k, d, &pc)) return -5;
    long long maxA = 0, aoff = 0;
    int maxB = 0;
    for (long long p = 0; p < n; p++) {
        const long long A = top_off[p + 1] - top_off[p], B = side_off[p + 1] - side_off[p];
        if (A > maxA) maxA = A;
        if (B > maxB) maxB = (int)B;
        arrow_off[p] = aoff;
        aoff += 128 * B;
    }
    arrow_off[n] = aoff;
    if (!nwb_bp_usable(pc, maxA)) return -6;
    std::vector<long long> fb((size_t)n + 1, -1);
    unsigned fbn = 0;
    NwbBpParams p;
    memset(&p, 0, sizeof(p));
    p.tops = (const uint8_t *)tops; p.top_off = top_off; p.sides = (const uint8_t *)sides; p.side_off = side_off;
    p.n_pairs = n; p.d = d; p.arrows = arrows; p.arrow_off = arrow_off; p.out_score = scores; p.out_branch = branches;
    p.fb_list = fb.data(); p.fb_count = &fbn; p.k2 = 2u; p.k4 = 4u;
    bool ok = false;
    const int nw = force_words > 0 ? force_words : nwb_bp_words(maxA);
    if (32 * nw < maxA) return -6;
    switch (pc.a_match) {
    case 1: ok = emu_bp_go<1>(pc.a_mis, nw, grid, warps, p); break;
    case 2: ok = emu_bp_go<2>(pc.a_mis, nw, grid, warps, p); break;
    case 3: ok = emu_bp_go<3>(pc.a_mis, nw, grid, warps, p); break;
    }
    if (!ok) return -6;
    if (n_fallback) *n_fallback = fbn;
    if (fbn) {
        NwbBatchParams bp;
        memset(&bp, 0, sizeof(bp));
        bp.tops = (const uint8_t *)tops; bp.top_off = top_off; bp.sides = (const uint8_t *)sides; bp.side_off = side_off;
        bp.n_pairs = n; bp.m = m; bp.k = k; bp.d = d; bp.max_B = maxB;
        bp.arrows = arrows; bp.arrow_off = arrow_off; bp.out_score = scores; bp.out_branch = branches;
        bp.bpitch = nwb_round_up((size_t)maxB + 1 + 64 + 256, 32);
        bp.pair_list = fb.data(); bp.pair_count = &fbn;
        emu_launch(grid, 32 * NWB_BATCH_WARPS, NWB_BATCH_SMEM_PER_WARP(maxB) * NWB_BATCH_WARPS, [&]() { nwb_batch_pk_kernel(bp, pc); });
    }
    return 0;
}
}
