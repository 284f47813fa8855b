// The count sweep's row step in isolation (one warp): cycles per step of CPL cells per lane.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o cnt cnt.cu && ./cnt
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define STEPS 4096
typedef unsigned long long u64;

template <int CPL>
__device__ __forceinline__ void row(unsigned x, u64 (&cnt)[CPL], u64 &left_above, u64 cl, u64 &send)
{
    u64 cd = left_above;
    left_above = cl;
#pragma unroll
    for (int k = 0; k < CPL; k++) {
        const unsigned f = x >> (4 * k);
        const u64 cu = cnt[k];
        const u64 n = ((f & 1u) ? cd : 0ull) + ((f & 2u) ? cl : 0ull) + ((f & 4u) ? cu : 0ull);
        cd = cu; cnt[k] = n; cl = n;
    }
    send = cl;
}

// MODE 0: two 32-bit shuffles (64-bit), lane 0 from smem; 1: no lane-0 smem; 2: only the low word shuffled; 3: no shuffle
template <int CPL, int MODE>
__global__ void k(u64 *out, long long *cyc, const unsigned *in)
{
    __shared__ u64 cstage[8];
    const int lane = threadIdx.x & 31;
    if (lane < 8) cstage[lane] = in[lane];
    __syncwarp();
    u64 cnt[CPL];
    for (int i = 0; i < CPL; i++) cnt[i] = 1;
    u64 send = 1, left_above = 1;
    long long t0 = clock64();
#pragma unroll 1
    for (int s = 0; s < STEPS; s += 8) {
        unsigned w[8];
#pragma unroll
        for (int t = 0; t < 8; t++) w[t] = in[64 + ((s + t + lane) & 255)];
#pragma unroll
        for (int t = 0; t < 8; t++) {
            u64 cl;
            if (MODE == 3) cl = send + 1;
            else if (MODE == 2) cl = (u64)__shfl_up_sync(0xffffffffu, (unsigned)send, 1) | (send & 0xffffffff00000000ull);
            else cl = __shfl_up_sync(0xffffffffu, send, 1);
            if (MODE == 0 && lane == 0) cl = cstage[t];
            row<CPL>(w[t], cnt, left_above, cl, send);
        }
    }
    long long t1 = clock64();
    u64 x = send;
    for (int i = 0; i < CPL; i++) x ^= cnt[i];
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main()
{
    u64 *out; unsigned *in; long long *cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8); cudaMalloc(&in, 4096);
    unsigned h[1024];
    for (int i = 0; i < 1024; i++) h[i] = (i * 2654435761u >> 5) | 0x11111111u;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
#define RUN(C, M) { for (int r = 0; r < 2; r++) k<C, M><<<1, 32>>>(out, cyc, in); long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); \
                    printf("CPL=%d mode %d  %7.1f cycles/step\n", C, M, (double)c / STEPS); }
    RUN(2, 0) RUN(2, 1) RUN(2, 2) RUN(2, 3) RUN(4, 0) RUN(8, 0) RUN(8, 1) RUN(8, 3)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
