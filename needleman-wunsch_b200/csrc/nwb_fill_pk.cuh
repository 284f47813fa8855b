/* nwb_fill_pk.cuh -- packed 16x2 difference kernel (placeholder until implemented). */
#pragma once
#include "nwb_device.cuh"
#define NWB_PK_WARPS 4
struct NwbPkConsts { int a_match, a_mis, c; };
static inline bool nwb_pk_supported(int, int, int, NwbPkConsts *) { return false; }
static inline int nwb_pk_choose_k(int, int, int) { return 1; }
#ifndef NWB_EMU
typedef int (*nwb_fail_fn)(cudaError_t, const char *);
static inline int nwb_pk_launch(const NwbStripParams &, const NwbPkConsts &, int, bool, int, cudaStream_t, nwb_fail_fn)
{
    return -5;
}
#endif
