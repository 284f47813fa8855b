/*
 * nwb_peak.cuh -- measurement aid: the INT/DPX issue rate of this GPU
 * ("R = INT/DPX thread-results per clk per SM -- measure on the
 * box with a microbenchmark of independent VIMNMX3/VIADDMNMX chains at full
 * occupancy").  Not on the fill path.
 *   mode 0: VIADDMNMX (s32)          1 result / thread-instruction
 *   mode 1: VIMNMX3 (s32)            1 result / thread-instruction
 *   mode 2: VIMNMX3.U16x2            1 instruction (2 packed results)
 *   mode 3: VIMNMX3 + IMAD pairs     ALU pipe + FMA pipe together
 */
#pragma once
#include "nwb_device.cuh"

#ifndef NWB_EMU
template <int MODE>
__global__ void __launch_bounds__(1024, 2) nwb_peak_kernel(int iters, int a, int b, int *sink, long long *cycles)
{
    int x0 = threadIdx.x, x1 = x0 + a, x2 = x0 ^ b, x3 = x0 - a, x4 = x0 + 7, x5 = x0 * 3, x6 = x0 - 11, x7 = x0 ^ 5;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0) {
                x0 = __viaddmax_s32(x0, a, b); x1 = __viaddmax_s32(x1, a, b); x2 = __viaddmax_s32(x2, a, b); x3 = __viaddmax_s32(x3, a, b);
                x4 = __viaddmax_s32(x4, a, b); x5 = __viaddmax_s32(x5, a, b); x6 = __viaddmax_s32(x6, a, b); x7 = __viaddmax_s32(x7, a, b);
            } else if (MODE == 1) {
                x0 = __vimax3_s32(x0, x1, a); x1 = __vimax3_s32(x1, x2, b); x2 = __vimax3_s32(x2, x3, a); x3 = __vimax3_s32(x3, x4, b);
                x4 = __vimax3_s32(x4, x5, a); x5 = __vimax3_s32(x5, x6, b); x6 = __vimax3_s32(x6, x7, a); x7 = __vimax3_s32(x7, x0, b);
            } else if (MODE == 2) {
                x0 = __vimin3_u16x2(x0, x1, a); x1 = __vimax3_u16x2(x1, x2, b); x2 = __vimin3_u16x2(x2, x3, a); x3 = __vimax3_u16x2(x3, x4, b);
                x4 = __vimin3_u16x2(x4, x5, a); x5 = __vimax3_u16x2(x5, x6, b); x6 = __vimin3_u16x2(x6, x7, a); x7 = __vimax3_u16x2(x7, x0, b);
            } else {
                x0 = __vimax3_s32(x0, x2, a); x1 = x1 * a + b; x2 = __vimax3_s32(x2, x4, b); x3 = x3 * a + b;
                x4 = __vimax3_s32(x4, x6, a); x5 = x5 * a + b; x6 = __vimax3_s32(x6, x0, b); x7 = x7 * a + b;
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if ((x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7) == 0x7fffffff) sink[0] = x0;
}
#endif
