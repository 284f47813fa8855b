/*
 * nwb_layout.h -- host-side geometry shared by the C-ABI implementation
 * (nwb_api.cu) and the test-only emulator harness: strip counts, pitches and
 * buffer sizes for one fill.  Plain C++ (no CUDA).
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#define NWB_KIND_I32 0
#define NWB_KIND_PK 1

struct NwbLayout {
    int A, B;
    int kind;        /* NWB_KIND_*                                   */
    int pk_k;        /* packed kernel: columns per half-lane (1..4)  */
    int pk_r;        /* packed kernel: rows per step (1 or 2)        */
    int strip_w;     /* interior columns per strip                   */
    int n_strips;
    size_t pitch;    /* arrow row pitch in bytes (multiple of 16)    */
    size_t spitch;   /* scores/cntmat elements per row               */
    size_t bpitch;   /* boundary stream elements per strip           */
};

static inline size_t nwb_round_up(size_t v, size_t q) { return (v + q - 1) / q * q; }

/* Geometry for `A x B` with strips of `strip_w` columns. */
static inline NwbLayout nwb_make_layout(int A, int B, int kind, int pk_k, int strip_w)
{
    NwbLayout L;
    L.A = A;
    L.B = B;
    L.kind = kind;
    L.pk_k = pk_k;
    L.pk_r = 1;
    L.strip_w = strip_w;
    L.n_strips = (A + strip_w - 1) / strip_w;
    if (L.n_strips < 1) L.n_strips = 1;
    L.pitch = nwb_round_up((size_t)L.n_strips * (size_t)strip_w / 2, 16);
    L.spitch = (size_t)L.n_strips * (size_t)strip_w;
    L.bpitch = nwb_round_up((size_t)B + 1 + 64 + 512, 32); /* + front/back padding of the packed kernel's streams */
    return L;
}

/* Strip group: rank r of `world` owns strips [begin, end) = r * ceil(n / world) ... (clamped to n).  The
 * reference's cyclic column sets (needleman-wunsch.c:568-571) interleave columns between threads; across
 * GPUs the strips are contiguous so that one boundary column per rank crosses NVLink. */
static inline void nwb_rank_strip_range(int n_strips, int rank, int world, int *begin, int *end)
{
    const int per = (n_strips + world - 1) / world;
    *begin = rank * per < n_strips ? rank * per : n_strips;
    *end = (rank + 1) * per < n_strips ? (rank + 1) * per : n_strips;
}

/* Batch of pairs: rank r takes the contiguous range [first, first + count); the first n % world ranks
 * take one pair more. */
static inline void nwb_rank_pair_range(long long n_pairs, int rank, int world, long long *first, long long *count)
{
    const long long q = n_pairs / world, r = n_pairs % world;
    *first = rank * q + (rank < r ? rank : r);
    *count = q + (rank < r ? 1 : 0);
}
