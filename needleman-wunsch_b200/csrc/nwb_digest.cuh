/*
 * nwb_digest.cuh -- on-device digests of the fill's results, for bit-exact parity
 * checks at sizes where the tables cannot be brought to the host (5 GB of arrow
 * codes at 100k x 100k, 33 GB for the 1M-pair batch).
 *
 * The digests are SUMS mod 2^64 of a mixing function over (position, value) pairs,
 * so they can be accumulated in any order, by any partition of the table (strips,
 * GPUs: rank digests simply add up).  The CPU oracle computes the same sums
 * (oracle/nw_oracle.h: nwo_mix64, nwo_result.arrow_digest / lastrow_count_digest /
 * lastcol_count_digest) from the reference's own definition of every cell
 * (needleman-wunsch.c:485-503 arrows, computation.c:223-260 count).
 *
 *   arrow digest  = sum over rows j = 1..B and 32-bit words w of the nibble table of
 *                   mix64(j << 32 | w, word(j,w) & 0x77777777, cells beyond column A zeroed)
 *   batch digests = sum over pairs p of mix64(first_pair + p, x_p), x_p = the pair's arrow
 *                   digest / optimal score / branch count / alignment count
 */
#pragma once
#include "nwb_device.cuh"

__host__ __device__ __forceinline__ unsigned long long nwb_mix64(unsigned long long pos, unsigned long long x)
{
    unsigned long long z = (pos + 1ull) * 0x9E3779B97F4A7C15ull + x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* the digest word of row j (1-based), word index w, for a table of A columns */
__device__ __forceinline__ unsigned long long nwb_arrow_word_term(const unsigned raw, const int j, const int w, const int A)
{
    unsigned x = raw & 0x77777777u;
    const int left = A - 8 * w; /* cells of this word inside the table */
    if (left < 8) x &= (left <= 0) ? 0u : ((1u << (4 * left)) - 1u);
    return nwb_mix64(((unsigned long long)(unsigned)j << 32) | (unsigned long long)(unsigned)w, x);
}

__device__ __forceinline__ unsigned long long nwb_warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NWB_FULL_MASK, v, o);
    return v;
}

/* Words [w_begin, w_end) of every row: one rank's strips, or the whole table. */
__global__ void __launch_bounds__(256) nwb_arrow_digest_kernel(const uint8_t *arrows, const size_t pitch, const int A, const int B,
                                                               const int w_begin, const int w_end, unsigned long long *out)
{
    const int nw = w_end - w_begin;
    unsigned long long acc = 0ull;
    if (nw > 0) {
        const long long total = (long long)nw * B;
        for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
            const int row = (int)(idx / nw), w = w_begin + (int)(idx - (long long)row * nw);
            const unsigned raw = *reinterpret_cast<const unsigned *>(arrows + (size_t)row * pitch + (size_t)w * 4);
            acc += nwb_arrow_word_term(raw, row + 1, w, A);
        }
    }
    acc = nwb_warp_sum_u64(acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

/* Batch: one warp per pair (grid-stride).  out[0..3] += mix64(first_pair + p, {arrow digest, score, branches, count}). */
__global__ void __launch_bounds__(256) nwb_batch_digest_kernel(const uint8_t *arrows, const long long *arrow_off,
                                                               const long long *top_off, const long long *side_off,
                                                               const long long n_pairs, const long long first_pair,
                                                               const int *score, const unsigned *branch,
                                                               const unsigned long long *count, unsigned long long *out)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned long long a0 = 0ull, a1 = 0ull, a2 = 0ull, a3 = 0ull;
    for (long long p = warp; p < n_pairs; p += nwarps) {
        const int A = (int)(top_off[p + 1] - top_off[p]), B = (int)(side_off[p + 1] - side_off[p]);
        const int ns = (A + 255) / 256 > 0 ? (A + 255) / 256 : 1;
        const size_t pitch = (size_t)ns * 128;
        const uint8_t *tab = arrows + arrow_off[p];
        const int nw = (A + 7) / 8;
        unsigned long long acc = 0ull;
        const long long total = (long long)nw * B;
        for (long long idx = lane; idx < total; idx += 32) {
            const int row = (int)(idx / nw), w = (int)(idx - (long long)row * nw);
            const unsigned raw = *reinterpret_cast<const unsigned *>(tab + (size_t)row * pitch + (size_t)w * 4);
            acc += nwb_arrow_word_term(raw, row + 1, w, A);
        }
        acc = nwb_warp_sum_u64(acc);
        if (lane == 0) {
            const unsigned long long g = (unsigned long long)(first_pair + p);
            a0 += nwb_mix64(g, acc);
            a1 += nwb_mix64(g, (unsigned long long)(long long)score[p]);
            if (branch) a2 += nwb_mix64(g, (unsigned long long)branch[p]);
            if (count) a3 += nwb_mix64(g, count[p]);
        }
    }
    if (lane == 0) {
        if (a0) atomicAdd(out + 0, a0);
        if (a1) atomicAdd(out + 1, a1);
        if (a2) atomicAdd(out + 2, a2);
        if (a3) atomicAdd(out + 3, a3);
    }
}
