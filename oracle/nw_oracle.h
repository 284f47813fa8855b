/*
 * nw_oracle.h -- CPU oracle for the Needleman-Wunsch score-table fill.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and only as the checker.
 *
 * It restates, in plain C, what the reference computes on its hot path:
 *   score_cell()              /root/reference/needleman-wunsch.c:418-510
 *   score_cell_column()       /root/reference/needleman-wunsch.c:523-543
 *   init_computation_tables() /root/reference/computation.c:75-125
 *   inc_branch_count()        /root/reference/walk-table.c:108-120
 * plus the optimal-alignment count that the reference obtains by enumeration
 * (needleman-wunsch.c:249-255, computation.c:223-235) restated as a 64-bit
 * path-count DP.
 *
 * Parity is PINNED: tests/test_oracle.py checks this file against
 *   - the README known answers (/root/reference/README:117-173),
 *   - the golden values SURVEY.md 8c recorded from the compiled reference,
 *   - the reference itself, compiled from /root/reference into oracle/_ref/
 *     (oracle/ref_harness.c) whenever /root/reference is present.
 *
 * Table orientation (reference naming): the TOP string s1 (length A) indexes
 * columns i in [0,A]; the SIDE string s2 (length B) indexes rows j in [0,B].
 */
#ifndef NW_ORACLE_H
#define NW_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* arrow code bits -- the same meaning as include/nwb.h */
#define NWO_DIAG 1u
#define NWO_LEFT 2u
#define NWO_UP 4u
#define NWO_MATCH 8u

typedef struct nwo_result {
    int32_t final_score;    /* cells[M-1][N-1].score                         */
    uint32_t branch_count;  /* # interior cells with >= 2 arrows (mod 2^32)  */
    int32_t greatest_abs;   /* max |score| over INTERIOR cells (tflag rule)  */
    uint32_t pad0;
    uint64_t table_hash;    /* FNV-1a-64 over (score, code) words, SURVEY 8c */
    uint64_t arrow_hash;    /* FNV-1a-64 over the 3-bit arrow codes only     */
    uint64_t count;         /* # optimal alignments mod 2^64                 */
    uint64_t count_hash;    /* FNV-1a-64 over every cell's count (lo,hi)     */
    uint64_t lastrow_count_hash; /* same, bottom row only (i = 0..A)         */
    uint64_t lastcol_count_hash; /* same, right column only (j = 0..B)       */
    /* Order-independent digests (sums mod 2^64 of nwo_mix64 terms): what the GPU
     * computes on device at full size, where a sequential FNV cannot be used.  */
    uint64_t arrow_digest;         /* sum over rows j=1..B, words w: mix64(j<<32|w, word(j,w));
                                    * word(j,w) = the 8 codes&7 of cells i=8w+1..8w+8 as nibbles
                                    * (nibble n <-> i=8w+1+n; cells beyond A are 0): the 32-bit
                                    * words of the include/nwb.h nibble table                 */
    uint64_t lastrow_count_digest; /* sum over i=1..A of mix64(i, cnt(i,B))  */
    uint64_t lastcol_count_digest; /* sum over j=1..B of mix64(j, cnt(A,j))  */
} nwo_result;

/* The digest's mixing function (SplitMix64 finaliser over a position/value pair). */
static inline uint64_t nwo_mix64(uint64_t pos, uint64_t x)
{
    uint64_t z = (pos + 1) * 0x9E3779B97F4A7C15ULL + x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/*
 * Fill the whole table.  Every output pointer may be NULL.
 *   scores : (B+1)*(A+1) int32, row-major [j][i], borders included
 *   codes  : (B+1)*(A+1) bytes, row-major [j][i], bits NWO_*; borders included
 *            (row 0: LEFT, col 0: UP, (0,0): 0; match bit only on interior)
 *   counts : (B+1)*(A+1) uint64, row-major, borders = 1
 *   packed : interior-only nibble table in the include/nwb.h layout:
 *            B rows of `pitch` bytes; cell (i,j), i,j>=1 lives in byte
 *            (j-1)*pitch + (i-1)/2, low nibble when (i-1) is even; the nibble
 *            is DIAG|LEFT|UP (bit 3 is always 0).
 * Returns 0, or -1 on bad arguments / allocation failure.
 */
int nwo_fill(const char *top, int A, const char *side, int B,
             int m, int k, int d,
             int32_t *scores, uint8_t *codes, uint64_t *counts,
             uint8_t *packed, size_t pitch,
             nwo_result *res);

/* SplitMix64 sequence generator of SURVEY.md 8d: n characters drawn iid
 * from alphabet[0..alen) with state seeded by `seed`. */
void nwo_generate(uint64_t seed, const char *alphabet, int alen, char *out, size_t n);

/*
 * Enumerate every optimal alignment from a codes[] table produced by nwo_fill,
 * in the reference's order (diag, then left, then up;
 * needleman-wunsch.c:209-331).  Calls cb(X, Y, len, user) for each alignment
 * (X, Y are the aligned strings in reading order, not NUL-terminated).
 * Stops after `limit` alignments if limit > 0.  Returns the number found.
 */
typedef void (*nwo_align_cb)(const char *X, const char *Y, int len, void *user);
uint64_t nwo_enumerate(const char *top, int A, const char *side, int B,
                       const uint8_t *codes, uint64_t limit,
                       nwo_align_cb cb, void *user);

#ifdef __cplusplus
}
#endif
#endif
