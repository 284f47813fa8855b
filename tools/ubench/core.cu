// The sweeping warp's step in isolation (one warp on an otherwise idle GPU): cycles per step of
// different formulations of the packed recurrence, K = 4 columns x R = 2 rows per lane and step.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o core core.cu && ./core
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define STEPS 4096

struct St { unsigned tpw[4], u[4], vlast[2], sp[2], send, nu; };

// V0: as nwb_hx_step (chars -> a by XOR + VIADDMNMX; z = max3; two subtractions)
// V1: a comes from registers (query profile), rest as V0
// V2: V1 + SWAR packing (Horner) as in the kernel
// V3: V0 without the second shuffle / SEL
// V4: chain through VIADDMNMX: z_k = max(z_{k-1} - u_{k-1}, max(a_k, u_k))
template <int V>
__device__ __forceinline__ void step(St &st, unsigned chars, unsigned bq, int t, int lane, unsigned TT1, unsigned AMIS,
                                     const uint4 &pa0, const uint4 &pa1, unsigned &acc)
{
    unsigned recv = __shfl_up_sync(0xffffffffu, st.send, 1);
    if (V != 3) {
        const unsigned b = __shfl_sync(0xffffffffu, bq, t);
        if (lane == 0) recv = b;
    }
    unsigned vL[2];
    vL[0] = __byte_perm(recv, st.vlast[0], 0x5410);
    vL[1] = __byte_perm(recv, st.vlast[1], 0x5432);
    if (V == 0 || V == 3) {
        st.sp[0] = __byte_perm(chars, st.sp[0], 0x5410);
        st.sp[1] = __byte_perm(chars, st.sp[1], 0x5432);
    }
#pragma unroll
    for (int r = 0; r < 2; r++) {
        unsigned v = vL[r];
        unsigned z[4], a[4];
        if (V == 1 || V == 2 || V == 4) {
            const uint4 &av = r ? pa1 : pa0;
            a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
        }
        if (V == 4) {
            unsigned zprev = 0, nuprev = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned m = __vmaxs2(a[k], st.u[k]);
                z[k] = (k == 0) ? __vimax3_s16x2(a[k], v, st.u[k]) : __viaddmax_s16x2(zprev, nuprev, m);
                const unsigned vin = (k == 0) ? v : (zprev + nuprev);
                nuprev = 0u - st.u[k];
                zprev = z[k];
                st.u[k] = z[k] - vin;
            }
            v = zprev + nuprev;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (V == 0 || V == 3) {
                    const unsigned nx = st.tpw[k] ^ st.sp[r];
                    a[k] = __viaddmax_s16x2(nx, TT1, AMIS);
                }
                z[k] = __vimax3_s16x2(a[k], v, st.u[k]);
                const unsigned un = z[k] - v;
                const unsigned vn = z[k] - st.u[k];
                st.u[k] = un;
                v = vn;
            }
        }
        st.vlast[r] = v;
        if (V == 2) {
            const unsigned Z4 = ((z[3] * 16u + z[2]) * 16u + z[1]) * 16u + z[0];
            const unsigned A4 = ((a[3] * 16u + a[2]) * 16u + a[1]) * 16u + a[0];
            const unsigned NU = ((st.u[3] * 16u + st.u[2]) * 16u + st.u[1]) * 16u + st.u[0];
            acc ^= ((A4 - Z4 + 0x88888888u) & 0x88888888u) | NU;
            acc += st.nu - Z4;
            st.nu = NU;
        }
    }
    st.send = __byte_perm(st.vlast[0], st.vlast[1], 0x7632);
}

template <int V>
__global__ void k(unsigned *out, long long *cyc, const unsigned *in)
{
    const int lane = threadIdx.x & 31;
    St st;
    for (int i = 0; i < 4; i++) { st.tpw[i] = in[i + lane]; st.u[i] = 0; }
    st.vlast[0] = st.vlast[1] = 0x7fff7fffu; st.sp[0] = st.sp[1] = 0xffffffffu; st.send = 0x7fff7fffu; st.nu = 0;
    const unsigned TT1 = in[40], AMIS = in[41];
    uint4 pa0 = make_uint4(in[42], in[43], in[44], in[45]), pa1 = make_uint4(in[46], in[47], in[48], in[49]);
    unsigned acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int s = 0; s < STEPS; s += 8) {
        unsigned bq = in[50 + (s & 63) + (lane & 7)];
        unsigned ch[8];
#pragma unroll
        for (int t = 0; t < 8; t++) ch[t] = in[128 + ((s + t) & 255)];
#pragma unroll
        for (int t = 0; t < 8; t++) {
            step<V>(st, ch[t], bq, t, lane, TT1, AMIS, pa0, pa1, acc);
            if (V == 1 || V == 2 || V == 4) { pa0.x ^= ch[t] & 2; pa1.y ^= ch[t] & 2; }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = st.send ^ st.u[0] ^ st.u[1] ^ st.u[2] ^ st.u[3] ^ acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main()
{
    unsigned *out, *in; long long *cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8); cudaMalloc(&in, 4096);
    unsigned h[1024];
    for (int i = 0; i < 1024; i++) h[i] = (i * 2654435761u >> 7) & 0x00030003u;
    h[40] = 0x00040004u; h[41] = 0x00010001u;
    for (int i = 42; i < 50; i++) h[i] = 0x00010003u;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const char *names[] = {"V0 kernel core (xor, viaddmnmx, vimnmx3, 2 sub)", "V1 a from registers", "V2 V1 + SWAR packing",
                           "V3 V0 without 2nd shuffle", "V4 VIADDMNMX chain"};
#define RUN(M) { for (int r = 0; r < 2; r++) k<M><<<1, 32>>>(out, cyc, in); long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); \
                 printf("%-52s %7.1f cycles/step\n", names[M], (double)c / STEPS); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
