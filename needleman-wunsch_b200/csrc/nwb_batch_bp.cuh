/*
 * nwb_batch_bp.cuh -- bit-parallel batch fill ("bp"): one THREAD per pair, a table row is a handful of
 * bit-vectors (64, 128 or 256 bits by the longest top string), one addition per difference level resolves a whole row.
 *
 * Same results as the reference's score_cell() (needleman-wunsch.c:418-510: the recurrence and "every tie
 * gets its arrow", :485-503) for top strings of up to 256 characters with at most five distinct letters,
 * when the per-cell differences are small (M = 2d + m <= 3: DNA 1/1/1 has M = 3).
 *
 * With r(i,j) = score(i,j) + d(i+j) (nwb_fill_pk.cuh) the differences u = r(i,j) - r(i-1,j) and
 * v = r(i,j) - r(i,j-1) lie in [0, M] and per cell
 *     z = max(a, vL, uU),  u = z - vL,  v = z - uU,       a = M on a match, N = 2d - k otherwise
 *     DIAG <=> z == a,  LEFT <=> u == 0,  UP <=> v == 0.
 * Along a row, v(i) = max(y(i), v(i-1) - uU(i)) with y = max(a - uU, 0).  For the level sets
 * V_k = [v >= k] (bit i-1 = column i) this reads
 *     V_k = S_k | ((V_k << 1) & P),   P = [uU == 0],
 *     S_k = [y >= k] | OR_{t>=1} ((V_{k+t} << 1) & [uU <= t]),
 * and "seeds S run through the runs of P" is ONE multi-word addition (the carry does the running, as in
 * Myers' bit-vector algorithm):  V_k = S_k | (((S_k & P') + P') ^ P'),  P' = P >> 1.  M additions per
 * row, levels from M down to 1; u of the new row = v + uU - vL in bit-sliced binary, the three arrow
 * planes are zero tests on the same vectors: 22 logic instructions, 4 shifts and 3 additions per 32 cells
 * for DNA 1/1/1 (nwb_bp_row), against ~200 instructions for the packed-difference kernels.
 *
 * Mapping: lane = pair.  A warp sweeps 32 pairs row by row; every lane keeps its row state (u as binary
 * bit-planes, 8 words each) in registers, the match vectors of its top string (one per letter) in shared
 * memory, and needs no shuffle and no carry from any other lane.  The arrow planes become the ABI's
 * 4-bit codes (include/nwb.h) through a byte -> 8-nibble table in shared memory (one copy per lane, so
 * no bank conflicts; at a 64 KB-aligned address, so that one PRMT forms a look-up address), and a row of the warp (32 pairs x 128 bytes) leaves through a swizzled staging
 * buffer so that every store instruction writes whole 128-byte lines.
 *
 * Pairs whose top string has more than five distinct letters are not computed here: they are put on a
 * list that nwb_batch_pk_kernel works off afterwards (nwb_batch_api.inl).  tools/bp_proto.py is the
 * executable statement of the formulation (Python integers as bit-vectors, checked against the oracle).
 */
#pragma once
#include "nwb_batch.cuh"

#define NWB_BP_NW 8     /* 32-bit words per row vector, at most: up to 256 columns (instantiated: 2, 4, 8) */
#define NWB_BP_MIN_AUTO 0  /* top strings up to this length keep the warp kernels unless asked (nwb_tune batch_bp = 1); measured:
                            * 32 x 32 pairs already run 7x faster here than in nwb_batch_bx_kernel */
#define NWB_BP_NSYM 5   /* letters with a match vector (DNA with N) */
#define NWB_BP_MAXM 3   /* instantiated difference ranges */
#define NWB_BP_WARPS 15  /* at most, per block (128 registers per thread; shared memory, see below) */
#define NWB_BP_SIDE_ROWS 32
#define NWB_BP_LUT_BYTES (256 * 256) /* entry x: 256 bytes = 32 lanes x {T[x]} then 32 lanes x {T[~x]} */
#define NWB_BP_PEQ_BYTES (NWB_BP_NSYM * NWB_BP_NW * 32 * 4)
#define NWB_BP_ZERO_BYTES (NWB_BP_NW * 32 * 4) /* one all-zero match vector per block: a side letter the top string does not have */
#define NWB_BP_STAGE_BYTES (32 * 128)
#define NWB_BP_SIDE_BYTES (NWB_BP_SIDE_ROWS * 32)
#define NWB_BP_META_BYTES (32 * 8)
#define NWB_BP_WARP_SMEM (NWB_BP_PEQ_BYTES + NWB_BP_STAGE_BYTES + NWB_BP_SIDE_BYTES + NWB_BP_META_BYTES)
/* The table sits at a 64 KB-aligned address of the shared window so that ONE PRMT forms a look-up address (byte of
 * the plane word into bits 8..15 of {table | lane * 4}).  Such a block asks for all of the SM's shared memory; the
 * warps' own areas (10.25 KB each) fill what lies in front of and behind the table: 15 warps fit whether the window of
 * the dynamic part starts at 0 or behind a reserved kilobyte. */
#define NWB_BP_SMEM_MAX 232448
#define NWB_BP_SMEM_BYTES(warps) ((size_t)NWB_BP_LUT_BYTES + NWB_BP_ZERO_BYTES + (size_t)(warps) * NWB_BP_WARP_SMEM) /* table not aligned */
/* the aligned table for launches that fill the SMs on their own; a chunk of a refill (a few warps per SM) takes the
 * other instantiation, whose blocks are small enough for two of them to share an SM */
#define NWB_BP_ALIGNED_MIN_WARPS 8

struct NwbBpParams {
    const uint8_t *tops;
    const long long *top_off;   /* n_pairs + 1 */
    const uint8_t *sides;
    const long long *side_off;  /* n_pairs + 1 */
    long long n_pairs;
    int d;
    uint8_t *arrows;            /* all pairs' nibble tables (pitch 128: one strip) */
    const long long *arrow_off; /* byte offset of pair p's table, a multiple of 128 */
    int *out_score;             /* [n_pairs] */
    unsigned *out_branch;       /* [n_pairs] or NULL */
    long long *fb_list;         /* pairs this kernel leaves to nwb_batch_pk_kernel (more than 5 letters) */
    unsigned *fb_count;
    unsigned k2, k4;            /* 2 and 4, as run-time values (see nwb_bp_rows) */
};

/* usable for this batch?  (m, k, d) inside the packed range, M = 2d + m <= 3, every top string <= 256 */
static inline bool nwb_bp_usable(const NwbPkConsts &pc, long long max_A)
{
    return pc.a_match >= 1 && pc.a_match <= NWB_BP_MAXM && pc.a_mis >= 0 && pc.a_mis <= pc.a_match && max_A <= 32 * NWB_BP_NW;
}

/* words per row vector for a batch whose longest top string has max_A letters */
static inline int nwb_bp_words(long long max_A) { return max_A <= 64 ? 2 : (max_A <= 128 ? 4 : 8); }

/* Warps per block: every group of 32 pairs costs the same, so the groups are worked off in rounds of
 * grid x warps; among the block sizes that need the fewest rounds take the smallest (the last round is the
 * fullest, and fewer resident warps leave each one more of the SM). */
static inline int nwb_bp_choose_warps(long long groups, int grid)
{
    if (groups <= 0 || grid <= 0) return NWB_BP_WARPS;
    const long long rounds = (groups + (long long)grid * NWB_BP_WARPS - 1) / ((long long)grid * NWB_BP_WARPS);
    int w = (int)((groups + rounds * grid - 1) / (rounds * grid));
    if (w < 4) w = 4;
    if (w > NWB_BP_WARPS) w = NWB_BP_WARPS;
    return w;
}

/* [x >= k] of a bit-sliced value (pl[0] = least significant plane); k is a compile-time constant after unrolling */
template <int NB, int NW>
__device__ __forceinline__ unsigned nwb_bp_ge(const unsigned (&pl)[NB][NW], const int k, const int w)
{
    if (k <= 0) return 0xFFFFFFFFu;
    if (k >= (1 << NB)) return 0u;
    unsigned res = 0xFFFFFFFFu;
#pragma unroll
    for (int t = 0; t < NB; t++) res = ((k >> t) & 1) ? (pl[t][w] & res) : (pl[t][w] | res);
    return res;
}

/* s = a + b over NW words (one carry chain) */
template <int NW>
__device__ __forceinline__ void nwb_bp_add(unsigned (&s)[NW], const unsigned (&a)[NW], const unsigned (&b)[NW])
{
#ifdef NWB_EMU
    unsigned long long c = 0ull;
    for (int w = 0; w < NW; w++) {
        c += (unsigned long long)a[w] + b[w];
        s[w] = (unsigned)c;
        c >>= 32;
    }
#else
    static_assert(NW == 2 || NW == 4 || NW == 8, "row vectors of 64, 128 or 256 bits");
    if constexpr (NW == 8) {
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
              "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    } else if constexpr (NW == 4) {
        asm("add.cc.u32 %0, %4, %8;\n\t"
            "addc.cc.u32 %1, %5, %9;\n\t"
            "addc.cc.u32 %2, %6, %10;\n\t"
            "addc.u32 %3, %7, %11;"
            : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
    } else {
        asm("add.cc.u32 %0, %2, %4;\n\t"
            "addc.u32 %1, %3, %5;"
            : "=r"(s[0]), "=r"(s[1])
            : "r"(a[0]), "r"(a[1]), "r"(b[0]), "r"(b[1]));
    }
#endif
}

/* binary planes of a level-set (thermometer) code th[1..M] */
template <int M, int NB, int NW>
__device__ __forceinline__ void nwb_bp_binary(unsigned (&out)[NB][NW], const unsigned (&th)[M + 2][NW])
{
#pragma unroll
    for (int w = 0; w < NW; w++) {
#pragma unroll
        for (int t = 0; t < NB; t++) {
            unsigned p = 0u;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int lo = (1 << t) * (2 * q + 1), hi = (1 << t) * (2 * q + 2);
                if (lo <= M) p |= (hi <= M) ? (th[lo][w] & ~th[hi][w]) : th[lo][w];
            }
            out[t][w] = p;
        }
    }
}

__device__ __forceinline__ unsigned nwb_bp_load_word(const uint8_t *s, const long long off, const bool aligned)
{
    if (aligned) return *reinterpret_cast<const unsigned *>(s + off);
    return (unsigned)s[off] | ((unsigned)s[off + 1] << 8) | ((unsigned)s[off + 2] << 16) | ((unsigned)s[off + 3] << 24);
}

/* d = f(a, b, c) bit by bit, f given by its truth table over NWB_LOP_A / _B / _C (one LOP3) */
#define NWB_LOP_A 0xF0
#define NWB_LOP_B 0xCC
#define NWB_LOP_C 0xAA
template <int LUT>
__device__ __forceinline__ unsigned nwb_lop3(const unsigned a, const unsigned b, const unsigned c)
{
#ifdef NWB_EMU
    unsigned r = 0u;
    for (int i = 0; i < 8; i++)
        if (((LUT & 0xFF) >> i) & 1) r |= ((i & 4) ? a : ~a) & ((i & 2) ? b : ~b) & ((i & 1) ? c : ~c);
    return r;
#else
    unsigned d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT & 0xFF));
    return d;
#endif
}

/* entry (byte q of x, or of ~x) of the lane's copy of the byte -> nibbles table.  lanetab: the shared-window address
 * of the lane's word of entry 0 (entry of ~x: + 128), bits 8..15 zero; the PRMT drops byte q of x into them. */
#ifdef NWB_EMU
typedef const unsigned *nwb_bp_tab;
template <bool INV, bool AL>
__device__ __forceinline__ unsigned nwb_bp_lut(nwb_bp_tab lanetab, const unsigned x, const int q)
{
    const unsigned idx = (x >> (8 * q)) & 0xFFu;
    return lanetab[idx * 64 + (INV ? 32 : 0)];
}
#else
typedef unsigned nwb_bp_tab;
template <bool INV, bool AL>
__device__ __forceinline__ unsigned nwb_bp_lut(nwb_bp_tab lanetab, const unsigned x, const int q)
{
    unsigned addr;
    if (AL) { /* the table is 64 KB-aligned: one PRMT */
        addr = __byte_perm(lanetab + (INV ? 128u : 0u), x, 0x3240u + ((unsigned)q << 4));
    } else {  /* anywhere: byte extraction (PRMT) and a multiply-add */
        addr = __byte_perm(x, 0u, 0x4440u + (unsigned)q) * 256u + (lanetab + (INV ? 128u : 0u));
    }
    unsigned v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
#endif

/* One row of one pair: E = match vector of the row's side letter, uu = u of the row above (bit-sliced; replaced by
 * this row's u).  Out: D = DIAG plane, Ln = NOT LEFT (u != 0), Un = NOT UP (v != 0). */
template <int M, int N, int NB, int NW>
__device__ __forceinline__ void nwb_bp_row(const unsigned (&E)[NW], unsigned (&uu)[NB][NW], unsigned (&D)[NW],
                                            unsigned (&Ln)[NW], unsigned (&Un)[NW])
{
    if constexpr (M == 3 && N == 1) {
        /* DNA 1/1/1, written out so that every intermediate is a function of three words, one LOP3 each (the
         * look-up tables are spelled out: the compiler does not find them from the generic expressions):
         * 22 logic instructions, 4 funnel shifts and 3 additions per 32 cells.  uU = (b1, b0). */
        constexpr int A_ = NWB_LOP_A, B_ = NWB_LOP_B, C_ = NWB_LOP_C;
        unsigned P[NW], Pp[NW], S[NW], T[NW], sum[NW], V3[NW], V2[NW], V1[NW], V3s[NW], V2s[NW], V1s[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) P[w] = nwb_lop3<~(A_ | B_)>(uu[1][w], uu[0][w], 0u);
#pragma unroll
        for (int w = 0; w < NW; w++) Pp[w] = __funnelshift_r(P[w], (w + 1 < NW) ? P[w + 1] : 0u, 1);
        /* level 3: y >= 3 <=> match and uU == 0 */
#pragma unroll
        for (int w = 0; w < NW; w++) {
            S[w] = nwb_lop3<A_ & B_>(E[w], P[w], 0u);
            T[w] = nwb_lop3<A_ & B_ & C_>(E[w], P[w], Pp[w]);
        }
        nwb_bp_add(sum, T, Pp);
#pragma unroll
        for (int w = 0; w < NW; w++) V3[w] = nwb_lop3<A_ | (B_ ^ C_)>(S[w], sum[w], Pp[w]);
#pragma unroll
        for (int w = 0; w < NW; w++) V3s[w] = (w > 0) ? __funnelshift_l(V3[w - 1], V3[w], 1) : (V3[0] << 1);
        /* level 2: (match or v(i-1) >= 3) and uU <= 1 */
#pragma unroll
        for (int w = 0; w < NW; w++) {
            S[w] = nwb_lop3<(A_ | B_) & ~C_>(E[w], V3s[w], uu[1][w]);
            T[w] = nwb_lop3<A_ & B_>(S[w], Pp[w], 0u);
        }
        nwb_bp_add(sum, T, Pp);
#pragma unroll
        for (int w = 0; w < NW; w++) V2[w] = nwb_lop3<A_ | (B_ ^ C_)>(S[w], sum[w], Pp[w]);
#pragma unroll
        for (int w = 0; w < NW; w++) V2s[w] = (w > 0) ? __funnelshift_l(V2[w - 1], V2[w], 1) : (V2[0] << 1);
        /* level 1: y >= 1, or v(i-1) >= 2 and uU <= 1, or v(i-1) >= 3 and uU <= 2 */
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const unsigned b1 = uu[1][w], b0 = uu[0][w];
            const unsigned y1 = nwb_lop3<(A_ & ~(B_ & C_)) | (~A_ & ~(B_ | C_))>(E[w], b1, b0);
            const unsigned g = nwb_lop3<A_ & ~B_ & C_>(b1, b0, V3s[w]);     /* uU == 2 and v(i-1) >= 3 */
            const unsigned h = nwb_lop3<A_ | (~B_ & C_)>(y1, b1, V2s[w]);   /* ... or uU <= 1 and v(i-1) >= 2 */
            S[w] = nwb_lop3<A_ | B_>(g, h, 0u);
            T[w] = nwb_lop3<(A_ | B_) & C_>(g, h, Pp[w]);
        }
        nwb_bp_add(sum, T, Pp);
#pragma unroll
        for (int w = 0; w < NW; w++) V1[w] = nwb_lop3<A_ | (B_ ^ C_)>(S[w], sum[w], Pp[w]);
#pragma unroll
        for (int w = 0; w < NW; w++) V1s[w] = (w > 0) ? __funnelshift_l(V1[w - 1], V1[w], 1) : (V1[0] << 1);
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const unsigned b1 = uu[1][w], b0 = uu[0][w];
            const unsigned v0 = nwb_lop3<A_ ^ B_ ^ C_>(V1[w], V2[w], V3[w]);          /* v = (V2, v0), vL = (V2s, l0) */
            const unsigned l0 = nwb_lop3<A_ ^ B_ ^ C_>(V1s[w], V2s[w], V3s[w]);
            const unsigned un0 = nwb_lop3<A_ ^ B_ ^ C_>(v0, l0, b0);                   /* u = v - vL + uU (mod 4) */
            const unsigned x1 = nwb_lop3<A_ ^ B_ ^ C_>(V2[w], V2s[w], b1);
            const unsigned x2 = nwb_lop3<(~A_ & B_) ^ ((A_ ^ B_) & C_)>(v0, l0, b0);   /* borrow of v - vL, carry of + uU */
            const unsigned un1 = nwb_lop3<A_ ^ B_>(x1, x2, 0u);
            const unsigned w1 = nwb_lop3<A_ ^ B_ ^ C_>(V2[w], b1, E[w]);               /* z = v + uU; DIAG <=> z == (match ? 3 : 1) */
            D[w] = nwb_lop3<(A_ ^ B_) & ~(C_ ^ (A_ & B_))>(v0, b0, w1);
            Ln[w] = nwb_lop3<A_ | B_>(un1, un0, 0u);
            Un[w] = V1[w];
            uu[0][w] = un0;
            uu[1][w] = un1;
        }
    } else {
        /* P' = [uU == 0] >> 1 */
        unsigned Pp[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const unsigned Pw = ~nwb_bp_ge<NB, NW>(uu, 1, w);
            const unsigned Pn = (w + 1 < NW) ? ~nwb_bp_ge<NB, NW>(uu, 1, w + 1) : 0u;
            Pp[w] = __funnelshift_r(Pw, Pn, 1);
        }
        unsigned V[M + 2][NW], Vs[M + 2][NW];
#pragma unroll
        for (int w = 0; w < NW; w++) { V[0][w] = 0xFFFFFFFFu; Vs[0][w] = 0xFFFFFFFFu; V[M + 1][w] = 0u; Vs[M + 1][w] = 0u; }
#pragma unroll
        for (int k = M; k >= 1; k--) {
            unsigned S[NW], T[NW], sum[NW];
#pragma unroll
            for (int w = 0; w < NW; w++) {
                unsigned y = E[w] & ~nwb_bp_ge<NB, NW>(uu, M - k + 1, w);
                if (k <= N) y |= ~E[w] & ~nwb_bp_ge<NB, NW>(uu, N - k + 1, w);
#pragma unroll
                for (int t = 1; t <= M - k; t++) y |= Vs[k + t][w] & ~nwb_bp_ge<NB, NW>(uu, t + 1, w);
                S[w] = y;
                T[w] = y & Pp[w];
            }
            nwb_bp_add(sum, T, Pp);
#pragma unroll
            for (int w = 0; w < NW; w++) V[k][w] = S[w] | (sum[w] ^ Pp[w]);
#pragma unroll
            for (int w = 0; w < NW; w++) Vs[k][w] = (w > 0) ? __funnelshift_l(V[k][w - 1], V[k][w], 1) : (V[k][0] << 1);
        }
        unsigned v[NB][NW], vl[NB][NW];
        nwb_bp_binary<M, NB, NW>(v, V);
        nwb_bp_binary<M, NB, NW>(vl, Vs);
#pragma unroll
        for (int w = 0; w < NW; w++) {
            /* u of this row = v + uU - vL (mod 2^NB: the result lies in [0, M]); z = v + uU */
            unsigned un[NB], z[NB];
            unsigned borrow = 0u, carry = 0u, zc = 0u;
#pragma unroll
            for (int t = 0; t < NB; t++) {
                const unsigned x = v[t][w] ^ vl[t][w] ^ borrow; /* v - vL */
                borrow = (~v[t][w] & (vl[t][w] | borrow)) | (vl[t][w] & borrow);
                un[t] = x ^ uu[t][w] ^ carry;
                carry = (x & uu[t][w]) | (carry & (x ^ uu[t][w]));
                z[t] = v[t][w] ^ uu[t][w] ^ zc;
                zc = (v[t][w] & uu[t][w]) | (zc & (v[t][w] ^ uu[t][w]));
            }
            unsigned eqM = 0xFFFFFFFFu, eqN = 0xFFFFFFFFu, nz = 0u;
#pragma unroll
            for (int t = 0; t < NB; t++) {
                eqM &= ((M >> t) & 1) ? z[t] : ~z[t];
                eqN &= ((N >> t) & 1) ? z[t] : ~z[t];
                nz |= un[t];
            }
            D[w] = (E[w] & eqM) | (~E[w] & eqN);
            Ln[w] = nz;
            Un[w] = V[1][w];
#pragma unroll
            for (int t = 0; t < NB; t++) uu[t][w] = un[t];
        }
    }
}

/* the distinct letters of a top string found so far */
struct NwbBpLetters {
    unsigned l[NWB_BP_NSYM];
    int nlet;
    bool over; /* a sixth letter: the pair is not for this kernel */
};

/* Up to four letters (the low `n` bytes of `word`, n >= 1) one at a time: new letters get the next free slot; bit
 * `bit0 + e` of acc[k] is set where letter e of the word is letter k.  The rare path of the match-vector build. */
__device__ __noinline__ void nwb_bp_letters_slow(NwbBpLetters &lt, const unsigned word, const int n, const int bit0,
                                                 unsigned (&acc)[NWB_BP_NSYM])
{
    for (int e = 0; e < 4 && e < n; e++) {
        const unsigned c = (word >> (8 * e)) & 0xFFu;
        int k = -1;
        for (int i = 0; i < NWB_BP_NSYM; i++)
            if (c == lt.l[i]) k = i;
        if (k < 0) {
            if (lt.nlet < NWB_BP_NSYM) {
                k = lt.nlet;
                lt.l[k] = c;
            } else {
                lt.over = true;
            }
            lt.nlet++;
        }
        if (k >= 0) acc[k] |= 1u << (bit0 + e);
    }
}

/* the arrow codes are written once and read much later, if at all: streaming stores (st.global.cs; 1.2 % on the
 * config 4 shard) */
__device__ __forceinline__ void nwb_bp_store(uint4 *p, const uint4 v)
{
#ifdef NWB_EMU
    *p = v;
#else
    __stcs(p, v);
#endif
}

/* The rows of one group of 32 pairs (lane = pair).  FULL: every pair of the group is 256 columns wide and all have
 * the same number of rows: no column masks, no per-pair row count at the stores, tables evenly spaced.  A
 * warp-uniform choice made outside the row loop, so that the common case carries no selects. */
template <int M, int N, int NB, int NW, bool FULL, bool AL>
__device__ __forceinline__ void nwb_bp_rows(const NwbBpParams &bp, uint4 *gdst, const unsigned *peq, uint4 *stage, uint8_t *side_sm,
                                             const uint2 *meta, const nwb_bp_tab mylut, const int lane, const int A, const int Brun,
                                             const int maxB, const long long s0, const bool s_al, const int nomatch,
                                             const unsigned l0, const unsigned l1, const unsigned l2, const unsigned l3, const unsigned l4,
                                             unsigned &branches, int &rsum)
{
    unsigned uu[NB][NW]; /* u of the row above, bit-sliced */
#pragma unroll
    for (int t = 0; t < NB; t++)
#pragma unroll
        for (int w = 0; w < NW; w++) uu[t][w] = 0u;
    /* Staging buffer: row = lane, 16-byte chunk w at position w ^ (lane & 7): conflict-free for the writers (a
     * quarter warp = 8 rows, 8 different positions) and for the readers (8 chunks of one row).  The multipliers of
     * the code-word combination are run-time values so that they stay multiply-adds (as constants they become
     * shifts and adds on the ALU pipe, which is the one that limits this kernel). */
    uint4 *st_w[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) st_w[w] = stage + lane * 8 + (w ^ (lane & 7));
    /* NW lanes store the row of one pair: lane = h * NW + ch, pair 32 / NW * i + h in round i */
    const int ch = lane & (NW - 1), h = lane / NW;
    const uint4 *st_r0 = stage + h * 8 + (ch ^ h), *st_r1 = stage + (h + 4) * 8 + (ch ^ (h + 4)); /* NW == 8: (L & 7) = h or h + 4 */
    const unsigned k2 = bp.k2, k4 = bp.k4;

#pragma unroll 1
    for (int j = 0; j < maxB; j++) {
        if ((j & (NWB_BP_SIDE_ROWS - 1)) == 0) {
            /* my next 64 side letters, one byte per row; only I read them back.  All 16 loads are issued before
             * the first store (one memory latency per 64 rows, not sixteen). */
#ifndef NWB_BP_STAGE_BATCH
#define NWB_BP_STAGE_BATCH 4 /* loads in flight; 16 measured 7 % slower (the row loop's schedule changes), 1 the same as 4 */
#endif
#pragma unroll 1
            for (int q0 = 0; q0 < NWB_BP_SIDE_ROWS / 4; q0 += NWB_BP_STAGE_BATCH) {
                unsigned word[NWB_BP_STAGE_BATCH];
#pragma unroll
                for (int q = 0; q < NWB_BP_STAGE_BATCH; q++) {
                    const int row = j + 4 * (q0 + q);
                    word[q] = (row < Brun) ? nwb_bp_load_word(bp.sides, s0 + row, s_al) : 0u;
                }
#pragma unroll
                for (int q = 0; q < NWB_BP_STAGE_BATCH; q++)
#pragma unroll
                    for (int e = 0; e < 4; e++) side_sm[(4 * (q0 + q) + e) * 32 + lane] = (uint8_t)(word[q] >> (8 * e));
            }
        }
        const unsigned c = side_sm[(j & (NWB_BP_SIDE_ROWS - 1)) * 32 + lane];
        /* word offset of the letter's match vector; a letter my top string does not have: the all-zero vector */
        const int voff = (c == l0) ? 0 : ((c == l1) ? NW * 32 : ((c == l2) ? 2 * NW * 32 : ((c == l3) ? 3 * NW * 32 : ((c == l4) ? 4 * NW * 32 : nomatch))));
        const unsigned *pe = peq + voff + lane;
        unsigned E[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) E[w] = pe[w * 32];

        /* one row: the arrow planes (LEFT and UP inverted) and u of this row */
        unsigned Dp[NW], Ln[NW], Un[NW];
        nwb_bp_row<M, N, NB, NW>(E, uu, Dp, Ln, Un);
        unsigned rowbr = 0u;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            if (!FULL) { /* columns beyond my top string: no arrows, no branch cells */
                const int left = A - 32 * w;
                const unsigned cm = (left >= 32) ? 0xFFFFFFFFu : ((left <= 0) ? 0u : ((1u << left) - 1u));
                Dp[w] &= cm; Ln[w] |= ~cm; Un[w] |= ~cm;
            }
            /* two or more arrows (walk-table.c:108-120): majority of DIAG, LEFT = ~Ln, UP = ~Un */
            rowbr += (unsigned)__popc(nwb_lop3<(NWB_LOP_A & ~NWB_LOP_B) | (NWB_LOP_A & ~NWB_LOP_C) | ~(NWB_LOP_B | NWB_LOP_C)>(Dp[w], Ln[w], Un[w]));
        }
        if (j < Brun) branches += rowbr;
        if (j == Brun - 1) { /* r(A,B) = sum of u(i,B) over my columns */
            int r = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) {
                const int left = A - 32 * w;
                const unsigned cm = (left >= 32) ? 0xFFFFFFFFu : ((left <= 0) ? 0u : ((1u << left) - 1u));
#pragma unroll
                for (int t = 0; t < NB; t++) r += __popc(uu[t][w] & cm) << t;
            }
            rsum = r;
        }
        /* planes -> 4-bit codes (DIAG | LEFT << 1 | UP << 2), my 128-byte row into the staging buffer.  With
         * T[x] = byte x spread to nibbles: code word = T[d] + 2 T[~ln] + 4 T[~un]; the byte comes out with one
         * PRMT, the table address (for ~x: counted down from entry 255) and the combination are multiply-adds
         * on the other pipe. */
#pragma unroll
        for (int w = 0; w < NW; w++) {
            unsigned o[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const unsigned td = nwb_bp_lut<false, AL>(mylut, Dp[w], q), tl = nwb_bp_lut<true, AL>(mylut, Ln[w], q), tu = nwb_bp_lut<true, AL>(mylut, Un[w], q);
                o[q] = tu * k4 + (tl * k2 + td);
            }
            *st_w[w] = make_uint4(o[0], o[1], o[2], o[3]);
        }
        __syncwarp();
        /* lanes h * NW .. h * NW + NW - 1 store the row of pair 32 / NW * i + h: whole 16 * NW-byte row pieces (NW = 8:
         * whole 128-byte lines) */
#pragma unroll
        for (int i = 0; i < NW; i++) {
            const int L = (32 / NW) * i + h;
            const uint4 val = (NW == 8) ? ((i & 1) ? st_r1 : st_r0)[(i >> 1) * 64] : stage[L * 8 + (ch ^ (L & 7))];
            if (FULL) {
                nwb_bp_store(gdst + ((unsigned)L * (unsigned)(8 * maxB) + (unsigned)(8 * j + ch)), val);
            } else {
                const uint2 mt = meta[L];
                if ((unsigned)j < mt.y) nwb_bp_store(gdst + (mt.x + (unsigned)(8 * j + ch)), val);
            }
        }
        __syncwarp();
    }
}

template <int M, int N, int NW, bool AL>
__global__ void __launch_bounds__(32 * NWB_BP_WARPS, 1) nwb_batch_bp_kernel(const NwbBpParams bp)
{
    constexpr int NB = (M >= 4) ? 3 : ((M >= 2) ? 2 : 1);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = (int)(blockDim.x >> 5);
    const long long nwarps = (long long)gridDim.x * wpb;
    const long long gwarp = (long long)blockIdx.x * wpb + warp;
    unsigned char *smem = NWB_SMEM_BASE();
#ifdef NWB_EMU
    unsigned char *lutp = smem;
    unsigned char *zerop = smem + NWB_BP_LUT_BYTES;
    unsigned char *mine = zerop + NWB_BP_ZERO_BYTES + (size_t)warp * NWB_BP_WARP_SMEM;
#else
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned before = AL ? ((sbase + 0xFFFFu) & ~0xFFFFu) - sbase : 0u; /* bytes in front of the (64 KB-aligned) table */
    unsigned char *lutp = smem + before;
    /* the block's zero vector and the first warps in front of the table, the others behind it */
    const bool zfront = before >= NWB_BP_ZERO_BYTES;
    const int nbefore = zfront ? (int)((before - NWB_BP_ZERO_BYTES) / NWB_BP_WARP_SMEM) : 0;
    unsigned char *zerop = zfront ? smem : lutp + NWB_BP_LUT_BYTES;
    unsigned char *behind = lutp + NWB_BP_LUT_BYTES + (zfront ? 0 : NWB_BP_ZERO_BYTES);
    unsigned char *mine = (warp < nbefore) ? smem + NWB_BP_ZERO_BYTES + (size_t)warp * NWB_BP_WARP_SMEM
                                           : behind + (size_t)(warp - nbefore) * NWB_BP_WARP_SMEM;
    if (AL && (size_t)(behind - smem) + (size_t)(wpb - (nbefore < wpb ? nbefore : wpb)) * NWB_BP_WARP_SMEM > NWB_BP_SMEM_MAX) __trap();
#endif
    unsigned *lut = reinterpret_cast<unsigned *>(lutp);
    unsigned *zero = reinterpret_cast<unsigned *>(zerop);                                /* [word][lane], all zero */
    unsigned *peq = reinterpret_cast<unsigned *>(mine);                                  /* [letter 0..4][word][lane] */
    uint4 *stage = reinterpret_cast<uint4 *>(mine + NWB_BP_PEQ_BYTES);                  /* [lane][8 swizzled 16-byte chunks] */
    uint8_t *side_sm = mine + NWB_BP_PEQ_BYTES + NWB_BP_STAGE_BYTES;                    /* [row & 31][lane] */
    uint2 *meta = reinterpret_cast<uint2 *>(mine + NWB_BP_PEQ_BYTES + NWB_BP_STAGE_BYTES + NWB_BP_SIDE_BYTES); /* {table row 0 / 128, rows} */

    /* byte -> 8 nibbles (bit i -> bit 4i), of x and of ~x, one copy per lane: the words of lane l sit in bank l */
    for (int x = warp; x < 256; x += wpb) {
        unsigned v = 0u;
#pragma unroll
        for (int i = 0; i < 8; i++) v |= ((unsigned)(x >> i) & 1u) << (4 * i);
        lut[x * 64 + lane] = v;
        lut[x * 64 + 32 + lane] = 0x11111111u - v;
    }
    if (warp == 0) {
#pragma unroll
        for (int w = 0; w < NW; w++) zero[w * 32 + lane] = 0u;
    }
    __syncthreads();
#ifdef NWB_EMU
    const nwb_bp_tab mylut = lut + lane;
#else
    const nwb_bp_tab mylut = sbase + before + 4u * (unsigned)lane;
#endif

    const long long groups = (bp.n_pairs + 31) / 32;
    for (long long g = gwarp; g < groups; g += nwarps) {
        const long long p = g * 32 + lane;
        const bool valid = p < bp.n_pairs;
        long long t0 = 0, s0 = 0;
        int A = 0, B = 0;
        if (valid) {
            t0 = bp.top_off[p];
            s0 = bp.side_off[p];
            A = (int)(bp.top_off[p + 1] - t0);
            B = (int)(bp.side_off[p + 1] - s0);
        }
        /* ---- match vectors of my top string: at most five distinct letters ---- */
        NwbBpLetters lt;
        lt.l[0] = 0x100u; lt.l[1] = 0x101u; lt.l[2] = 0x102u; lt.l[3] = 0x103u; lt.l[4] = 0x104u; /* no byte equals an unassigned letter */
        lt.nlet = 0;
        lt.over = false;
        const bool t_al = __all_sync(NWB_FULL_MASK, (t0 & 3) == 0);
        const bool t_al16 = __all_sync(NWB_FULL_MASK, (t0 & 15) == 0);
        int maxA = A;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int x = __shfl_xor_sync(NWB_FULL_MASK, maxA, o);
            maxA = x > maxA ? x : maxA;
        }
#pragma unroll 1
        for (int w = 0; w < NW; w++) {
            unsigned acc[NWB_BP_NSYM];
#pragma unroll
            for (int k = 0; k < NWB_BP_NSYM; k++) acc[k] = 0u;
            if (32 * w < maxA) {
                /* the 32 letters of this word: all loads first */
                unsigned wd[8];
                if (t_al16) {
#pragma unroll
                    for (int hq = 0; hq < 2; hq++) {
                        const int col = 32 * w + 16 * hq;
                        uint4 v4 = make_uint4(0u, 0u, 0u, 0u);
                        if (col < A) v4 = *reinterpret_cast<const uint4 *>(bp.tops + t0 + col);
                        wd[4 * hq + 0] = v4.x; wd[4 * hq + 1] = v4.y; wd[4 * hq + 2] = v4.z; wd[4 * hq + 3] = v4.w;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const int col = 32 * w + 4 * q;
                        wd[q] = (col < A) ? nwb_bp_load_word(bp.tops, t0 + col, t_al) : 0u;
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    /* four letters at a time: per known letter the bytes of the word that equal it (exact zero-byte
                     * test of word ^ letter x 0x01010101), gathered into four bits by a multiplication.  A word with a
                     * letter not seen before (the first few of a string) or with its end in it goes letter by letter. */
                    const int col = 32 * w + 4 * q;
                    const unsigned word = wd[q];
                    unsigned z[NWB_BP_NSYM], zall = 0u;
#pragma unroll
                    for (int k = 0; k < NWB_BP_NSYM; k++) {
                        const unsigned t = word ^ (lt.l[k] * 0x01010101u);
                        z[k] = (k < lt.nlet) ? (~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u) : 0u; /* slots in use only */
                        zall |= z[k];
                    }
                    if (col + 4 <= A && zall == 0x80808080u) {
#pragma unroll
                        for (int k = 0; k < NWB_BP_NSYM; k++)
                            acc[k] |= ((((z[k] >> 7) * 0x00204081u) >> 21) & 0xFu) << (4 * q);
                    } else if (col < A) {
                        nwb_bp_letters_slow(lt, word, A - col, 4 * q, acc);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < NWB_BP_NSYM; k++) peq[(k * NW + w) * 32 + lane] = acc[k];
        }
        const unsigned l0 = lt.l[0], l1 = lt.l[1], l2 = lt.l[2], l3 = lt.l[3], l4 = lt.l[4];
        const bool over = lt.over;
        if (valid && over) { /* not mine: the general batch kernel takes this pair */
            const unsigned pos = atomicAdd(bp.fb_count, 1u);
            bp.fb_list[pos] = p;
        }
        if (valid && !over && (A == 0 || B == 0)) { /* borders only */
            bp.out_score[p] = (A == 0) ? -B * bp.d : -A * bp.d;
            if (bp.out_branch) bp.out_branch[p] = 0u;
        }
        const int Brun = (valid && !over && A > 0) ? B : 0; /* rows this lane computes and stores */
        int maxB = Brun;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int x = __shfl_xor_sync(NWB_FULL_MASK, maxB, o);
            maxB = x > maxB ? x : maxB;
        }
        const bool s_al = __all_sync(NWB_FULL_MASK, (s0 & 3) == 0);
        __syncwarp();

        unsigned branches = 0u;
        int rsum = 0;
        /* tables of a group lie within 2^32 16-byte chunks of the first one (B <= 60000): 32-bit chunk indices */
        const long long off0 = __shfl_sync(NWB_FULL_MASK, valid ? bp.arrow_off[p] : 0ll, 0);
        /* one shape, 256 columns, tables one after the other */
        const bool full = __all_sync(NWB_FULL_MASK, A == 32 * NW && Brun == maxB && Brun > 0 &&
                                                        bp.arrow_off[valid ? p : 0] - off0 == (long long)lane * 128 * maxB);
        meta[lane] = make_uint2(valid ? (unsigned)((bp.arrow_off[p] - off0) >> 4) : 0u, (unsigned)Brun);
        __syncwarp();
        uint4 *gdst = reinterpret_cast<uint4 *>(bp.arrows + off0);
        if (full) nwb_bp_rows<M, N, NB, NW, true, AL>(bp, gdst, peq, stage, side_sm, meta, mylut, lane, A, Brun, maxB, s0, s_al, (int)(zero - peq), l0, l1, l2, l3, l4, branches, rsum);
        else nwb_bp_rows<M, N, NB, NW, false, AL>(bp, gdst, peq, stage, side_sm, meta, mylut, lane, A, Brun, maxB, s0, s_al, (int)(zero - peq), l0, l1, l2, l3, l4, branches, rsum);
        if (Brun > 0) {
            bp.out_score[p] = rsum - bp.d * (A + B);
            if (bp.out_branch) bp.out_branch[p] = branches;
        }
        __syncwarp();
    }
}
