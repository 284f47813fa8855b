// Issue-throughput micro-benchmarks: one warp per SM sub-partition (block of 128) or two (256),
// 8 independent dependency chains per thread, cycles per warp-instruction on one SMSP.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tput tput.cu && ./tput
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 2048
#define NCH 8
template <int MODE>
__device__ __forceinline__ unsigned op(unsigned x, unsigned y, unsigned c)
{
    unsigned r;
    if (MODE == 0) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(c));
    if (MODE == 1) asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(c));
    if (MODE == 2) { unsigned long long w; asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w) : "r"(x), "r"(y), "l"((unsigned long long)c)); r = (unsigned)(w >> 32) ^ (unsigned)w; }
    if (MODE == 3) asm volatile("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(x), "r"(y), "r"(c));
    if (MODE == 4) asm volatile("shf.r.wrap.b32 %0, %1, %2, 3;" : "=r"(r) : "r"(x), "r"(y));
    if (MODE == 5) asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(c));
    if (MODE == 6) asm volatile("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(r) : "r"(x), "r"(y));
    if (MODE == 7) asm volatile("popc.b32 %0, %1;" : "=r"(r) : "r"(x));
    if (MODE == 8) r = __vminu2(x, y);
    if (MODE == 9) r = __vimax3_s16x2(x, y, c);
    if (MODE == 10) r = __viaddmax_s16x2(x, y, c);
    if (MODE == 11) asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(c));
    if (MODE == 12) { float f; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(f) : "f"(__uint_as_float(x)), "f"(__uint_as_float(y)), "f"(__uint_as_float(c))); r = __float_as_uint(f); }
    if (MODE == 13) asm volatile("shr.u32 %0, %1, 3;" : "=r"(r) : "r"(x));
    if (MODE == 14) asm volatile("sub.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(c));
    if (MODE == 15) asm volatile("shl.b32 %0, %1, 4;" : "=r"(r) : "r"(x));
    return r;
}
template <int MODE, int MODE2>
__global__ void k(unsigned *out, long long *cyc, unsigned seed)
{
    unsigned v[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) v[i] = seed * (i + 3) + threadIdx.x;
    unsigned y = seed | 1, c = seed + 7;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            v[i] = op<MODE>(v[i], y, c);
            if (MODE2 >= 0) v[i] = op<(MODE2 >= 0 ? MODE2 : 0)>(v[i], y, c);
        }
    }
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s ^= v[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
    unsigned *out; long long *cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    const char *names[] = {"IMAD", "IMAD.HI.U32", "IMAD.WIDE.U32", "LOP3", "SHF.R(funnel)", "IADD(add)", "PRMT", "POPC", "VIMNMX.U16x2", "VIMNMX3.S16x2",
                           "VIADDMNMX.S16x2", "HFMA2", "FFMA", "SHR", "SUB", "SHL"};
#define RUN1(M, W) { for (int r = 0; r < 2; r++) k<M, -1><<<1, 128 * W>>>(out, cyc, 12345u); long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
                     printf("%-18s warps/SMSP=%d  %6.2f cyc/warp-instr/SMSP\n", names[M], W, (double)h / (ITER * NCH * W)); }
#define RUN2(M, M2, W) { for (int r = 0; r < 2; r++) k<M, M2><<<1, 128 * W>>>(out, cyc, 12345u); long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
                     printf("%-14s+%-14s warps/SMSP=%d  %6.2f cyc/warp-instr/SMSP (dependent pair)\n", names[M], names[M2], W, (double)h / (ITER * NCH * W * 2)); }
    RUN1(0, 1) RUN1(0, 2) RUN1(1, 1) RUN1(1, 2) RUN1(2, 1) RUN1(2, 2) RUN1(3, 1) RUN1(3, 2) RUN1(4, 1) RUN1(5, 1) RUN1(5, 2) RUN1(6, 1) RUN1(7, 1) RUN1(7, 2)
    RUN1(8, 1) RUN1(8, 2) RUN1(9, 1) RUN1(9, 2) RUN1(10, 1) RUN1(11, 1) RUN1(11, 2) RUN1(12, 1) RUN1(12, 2) RUN1(13, 1) RUN1(14, 1) RUN1(15, 1)
    RUN2(9, 0, 1) RUN2(9, 0, 2) RUN2(3, 0, 1) RUN2(3, 0, 2) RUN2(9, 3, 1) RUN2(9, 3, 2) RUN2(3, 1, 1) RUN2(3, 1, 2) RUN2(9, 12, 2) RUN2(9, 11, 2)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
