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
