// Latency micro-benchmarks for the packed fill's loop-carried chain (one warp, one block).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lat lat.cu && ./lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N 4096
__device__ __forceinline__ unsigned vimax3(unsigned a, unsigned b, unsigned c) { return __vimax3_s16x2(a, b, c); }

template <int MODE>
__global__ void k(unsigned *out, long long *cyc, unsigned seed, unsigned *sink)
{
    const int lane = threadIdx.x & 31;
    unsigned x = seed + lane, y = seed * 3 + 1, u = seed ^ 5;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N / 8; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (MODE == 0) x = __shfl_up_sync(0xffffffffu, x, 1);
            if (MODE == 1) x = vimax3(x, y, u);
            if (MODE == 2) { unsigned z = vimax3(x, y, u); x = z - u; }                  // max3 -> sub
            if (MODE == 3) { unsigned r = __shfl_up_sync(0xffffffffu, x, 1); unsigned z = vimax3(r, y, u); x = z - u; }
            if (MODE == 4) { unsigned r = __shfl_up_sync(0xffffffffu, x, 1); if (lane == 0) r = y; r = __byte_perm(r, u, 0x5432);
                             unsigned z = vimax3(r, y, u); x = z - u; }
            if (MODE == 5) { unsigned r = __shfl_up_sync(0xffffffffu, x, 1); unsigned z = vimax3(r, y, u); x = z - u;
                             if (lane == 31) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(sink + (i * 8 + j)), "r"(x)); }
            if (MODE == 6) { unsigned r = __shfl_up_sync(0xffffffffu, x, 1); unsigned z = vimax3(r, y, u); x = z - u;
                             asm volatile("st.shared.u32 [%0], %1;" ::"r"((unsigned)(lane * 4 + ((i * 8 + j) & 127) * 128)), "r"(x)); }
            if (MODE == 7) { unsigned z = vimax3(x, y, u); asm("sub.u32 %0, %1, %2;" : "=r"(x) : "r"(z), "r"(u)); x = __vminu2(x, 0x7fff7fffu); } // 3 ALU-ish
            if (MODE == 8) x = __byte_perm(x, y, 0x5410 + (j & 1));
            if (MODE == 9) { unsigned r = __shfl_up_sync(0xffffffffu, x, 1); unsigned b = __shfl_sync(0xffffffffu, y, j); if (lane == 0) r = b; unsigned z = vimax3(r, y, u); x = z - u; }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main()
{
    unsigned *out, *sink; long long *cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8); cudaMalloc(&sink, N * 4 + 64);
    const char *names[] = {"SHFL.UP chain", "VIMNMX3.S16x2 chain", "VIMNMX3 -> sub", "SHFL -> VIMNMX3 -> sub",
                           "SHFL -> SEL -> PRMT -> VIMNMX3 -> sub", "mode 3 + lane-31 st.relaxed.sys", "mode 3 + STS",
                           "VIMNMX3 -> IADD -> VMIN", "PRMT chain", "mode3 + second SHFL.IDX + SEL"};
#define RUN(M) { for (int r = 0; r < 2; r++) k<M><<<1, 32>>>(out, cyc, 12345u, sink); long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
                 printf("mode %d  %-40s %7.2f cycles/iter\n", M, names[M], (double)h / N); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
