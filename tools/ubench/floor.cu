// The DEPENDENCY FLOOR of the strip wavefront: how fast can one warp run the step of nwb_fill_hx.cuh when the step
// contains nothing but its dependent chain?  A step of K = 4 columns x R = 2 rows per half-lane is, on the critical
// path: SHFL.UP (the neighbour's v) -> PRMT -> (K + R - 1) = 5 cells of {VIMNMX3.S16x2 -> IMAD.IADD} -> PRMT -> next
// SHFL.UP.  Everything else of the real step (match terms, the other 11 max3 of the 2 x 4 x 2 block, packing, ring,
// stream stores) is off the chain and could in principle be done by other warps.
//
//   F0  the chain alone (1 shuffle + 2 PRMT + 5 x {max3, sub})            -> cycles per step = latency floor of K=4,R=2
//   F1  the chain of a K = 1, R = 1 step (1 shuffle + 1 x {max3, sub})    -> cycles per anti-diagonal of ONE cell
//   F2  F0 with the full recurrence of the step (16 max3 + 32 sub + match terms), no packing, no stores
//
// A 100,000 x 100,000 table swept in 256-column strips has 50,000 + 72 * 391 = 78,152 dependent steps
// (DESIGN 5); floor time = steps x cycles(F0) / f.  With one cell per step (F1) the chain is 200,000 anti-diagonals.
// Runs ONE warp per SM sub-partition on every SM (the occupancy of the real sweeping warps); the chain is latency-bound,
// so the number is the same for one warp alone.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o floor floor.cu && ./floor
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define STEPS 65536

template <int V>
__global__ void k(unsigned *out, long long *cyc, const unsigned *in)
{
    const int lane = threadIdx.x & 31;
    unsigned send = in[lane], vlast0 = in[32 + lane], vlast1 = in[64 + lane];
    unsigned a = in[96 + (lane & 3)], u0 = in[100], u1 = in[101], u2 = in[102], u3 = in[103], u4 = in[104];
    unsigned tp[4] = {in[110 + lane], in[142 + lane], in[174 + lane], in[206 + lane]};
    const unsigned TT1 = in[40], AMIS = in[41];
    unsigned acc = 0;
    long long t0 = clock64();
#pragma unroll 8
    for (int s = 0; s < STEPS; s++) {
        const unsigned recv = __shfl_up_sync(0xffffffffu, send, 1);
        if (V == 1) {
            const unsigned z = __vimax3_s16x2(a, recv, u0);
            send = z - u0;
            u0 = z - recv;
        } else {
            unsigned v = __byte_perm(recv, vlast0, 0x5410);
            if (V == 0) {
                // the 5 cells of the step's longest dependency path: (r0,k0) (r0,k1) (r0,k2) (r0,k3) (r1,k3)
                unsigned z;
                unsigned uo;
                z = __vimax3_s16x2(a, v, u0); uo = u0; u0 = z - v; v = z - uo;
                z = __vimax3_s16x2(a, v, u1); uo = u1; u1 = z - v; v = z - uo;
                z = __vimax3_s16x2(a, v, u2); uo = u2; u2 = z - v; v = z - uo;
                z = __vimax3_s16x2(a, v, u3); uo = u3; u3 = z - v; v = z - uo;
                vlast0 = v;
                /* sub-row 1's last cell needs sub-row 0's last u (just computed) and sub-row 1's own v */
                z = __vimax3_s16x2(a, vlast1, u3); u4 = z - vlast1; vlast1 = z - u3;
                acc ^= u4;
            } else {
                // the whole recurrence of the step, as in nwb_hx_step, without packing and stores
                unsigned vv[2] = {v, __byte_perm(recv, vlast1, 0x5432)};
                unsigned uu[4] = {u0, u1, u2, u3};
                const unsigned sp[2] = {in[300 + (s & 63)], in[364 + (s & 63)]};
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    unsigned w = vv[r];
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        const unsigned nx = tp[kk] ^ sp[r];
                        const unsigned aa = __viaddmax_s16x2(nx, TT1, AMIS);
                        const unsigned z = __vimax3_s16x2(aa, w, uu[kk]);
                        const unsigned un = z - w, vn = z - uu[kk];
                        uu[kk] = un;
                        w = vn;
                    }
                    vv[r] = w;
                }
                u0 = uu[0]; u1 = uu[1]; u2 = uu[2]; u3 = uu[3];
                vlast0 = vv[0]; vlast1 = vv[1];
            }
            send = __byte_perm(vlast0, vlast1, 0x7632);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = send ^ u0 ^ u1 ^ u2 ^ u3 ^ u4 ^ acc ^ vlast0 ^ vlast1;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int blocks = prop.multiProcessorCount, threads = 128; // one warp per SM sub-partition
    unsigned *out, *in; long long *cyc;
    cudaMalloc(&out, (size_t)blocks * threads * 4); cudaMalloc(&cyc, 8); cudaMalloc(&in, 4096);
    unsigned h[1024];
    for (int i = 0; i < 1024; i++) h[i] = (i * 2654435761u >> 7) & 0x00030003u;
    h[40] = 0x00040004u; h[41] = 0x00010001u;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const double f_ghz = prop.clockRate * 1e-6;
    const char *names[] = {"F0 chain of a K=4,R=2 step (SHFL + 2 PRMT + 5 x {VIMNMX3, IADD})", "F1 chain of one cell (SHFL + VIMNMX3 + IADD)",
                           "F2 F0 + the step's whole recurrence (no packing, no stores)"};
    double c[3];
#define RUN(M) { for (int r = 0; r < 3; r++) k<M><<<blocks, threads>>>(out, cyc, in); long long cc; cudaMemcpy(&cc, cyc, 8, cudaMemcpyDeviceToHost); \
                 c[M] = (double)cc / STEPS; printf("%-70s %7.1f cycles/step\n", names[M], c[M]); }
    RUN(0) RUN(1) RUN(2)
    const double steps = 50000.0 + 72.0 * 391.0;
    printf("SM clock %.0f MHz (cudaDeviceProp.clockRate)\n", f_ghz * 1e3);
    printf("100k x 100k, 256-column strips: %.0f dependent steps x %.1f cycles = %.2f ms at %.3f GHz (floor of the K=4,R=2 strip "
           "wavefront on ANY number of GPUs; nwb_fill_hx_kernel runs the same steps at ~176 cycles)\n",
           steps, c[0], steps * c[0] / (f_ghz * 1e6), f_ghz);
    printf("with the step's whole recurrence in the sweeping warp (F2): %.2f ms\n", steps * c[2] / (f_ghz * 1e6));
    printf("one cell per step (F1): 200,000 anti-diagonals x %.1f cycles = %.2f ms\n", c[1], 200000.0 * c[1] / (f_ghz * 1e6));
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
