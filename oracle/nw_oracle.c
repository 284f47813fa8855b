/*
 * nw_oracle.c -- CPU oracle for the Needleman-Wunsch score-table fill.
 * TEST INFRASTRUCTURE ONLY; see nw_oracle.h for what it restates and how its
 * parity is pinned.  Plain C99, rolling two rows, O(A) memory unless full
 * tables are requested.
 */
#include "nw_oracle.h"

#include <stdlib.h>
#include <string.h>

#define FNV_OFFSET 0xcbf29ce484222325ULL
#define FNV_PRIME 0x100000001b3ULL

static inline uint64_t fnv_word(uint64_t h, uint32_t w)
{
    h ^= (uint64_t)w;
    h *= FNV_PRIME;
    return h;
}

static inline uint64_t fnv_u64(uint64_t h, uint64_t v)
{
    h = fnv_word(h, (uint32_t)v);
    h = fnv_word(h, (uint32_t)(v >> 32));
    return h;
}

/* two's-complement wrap-around int add/sub (the reference's `int` arithmetic
 * never overflows on sane inputs; we define the overflow case instead of
 * inheriting C's undefined behaviour) */
static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }

/* reference: max3(), needleman-wunsch.c:395-404 */
static inline int32_t max3(int32_t a, int32_t b, int32_t c)
{
    int32_t m = a;
    if (m < b) m = b;
    if (m < c) m = c;
    return m;
}

static inline int32_t iabs32(int32_t v) { return v < 0 ? wsub(0, v) : v; }

int nwo_fill(const char *top, int A, const char *side, int B,
             int m, int k, int d,
             int32_t *scores, uint8_t *codes, uint64_t *counts,
             uint8_t *packed, size_t pitch,
             nwo_result *res)
{
    if (A < 0 || B < 0 || (A > 0 && !top) || (B > 0 && !side)) return -1;
    if (packed && pitch < (size_t)((A + 1) / 2)) return -1;
    const size_t M = (size_t)A + 1; /* columns */
    const size_t N = (size_t)B + 1; /* rows    */

    int32_t *prev = (int32_t *)malloc(M * sizeof(int32_t));
    int32_t *cur = (int32_t *)malloc(M * sizeof(int32_t));
    uint64_t *cprev = (uint64_t *)malloc(M * sizeof(uint64_t));
    uint64_t *ccur = (uint64_t *)malloc(M * sizeof(uint64_t));
    if (!prev || !cur || !cprev || !ccur) {
        free(prev); free(cur); free(cprev); free(ccur);
        return -1;
    }

    uint64_t th = FNV_OFFSET, ah = FNV_OFFSET, ch = FNV_OFFSET;
    uint64_t lrh = FNV_OFFSET, lch = FNV_OFFSET;
    uint64_t adig = 0, lrdig = 0, lcdig = 0;
    uint32_t branches = 0;
    int32_t gabs = 0;

    /* row 0: init_computation_tables(), computation.c:97-114.
     * (0,0): score 0, no arrows; (i,0): score i*(-d), LEFT. */
    for (size_t i = 0; i < M; i++) {
        prev[i] = wmul((int32_t)i, wsub(0, d));
        cprev[i] = 1;
        uint8_t code = (i == 0) ? 0 : NWO_LEFT;
        if (scores) scores[i] = prev[i];
        if (codes) codes[i] = code;
        if (counts) counts[i] = 1;
        th = fnv_word(th, (uint32_t)prev[i]);
        th = fnv_word(th, code);
        ah = fnv_word(ah, code & 7u);
        ch = fnv_u64(ch, 1);
        if (B == 0) lrh = fnv_u64(lrh, 1);
    }
    if (A == 0 || B == 0) lch = fnv_u64(lch, 1);
    else lch = fnv_u64(lch, cprev[A]);

    if (packed && B > 0) memset(packed, 0, pitch * (size_t)B);

    for (size_t j = 1; j < N; j++) {
        /* column 0: score j*(-d), UP (computation.c:116-124) */
        cur[0] = wmul((int32_t)j, wsub(0, d));
        ccur[0] = 1;
        if (scores) scores[j * M] = cur[0];
        if (codes) codes[j * M] = NWO_UP;
        if (counts) counts[j * M] = 1;
        th = fnv_word(th, (uint32_t)cur[0]);
        th = fnv_word(th, NWO_UP);
        ah = fnv_word(ah, NWO_UP);
        ch = fnv_u64(ch, 1);
        if (j == N - 1) lrh = fnv_u64(lrh, 1);

        const char sc = side[j - 1];
        uint8_t *prow = packed ? packed + (j - 1) * pitch : NULL;
        uint32_t dword = 0; /* the nibble-table word being assembled for the digest */
        for (size_t i = 1; i < M; i++) {
            /* score_cell(), needleman-wunsch.c:418-510 */
            const int32_t up = wsub(prev[i], d);
            const int32_t left = wsub(cur[i - 1], d);
            int32_t diag;
            uint8_t code = 0;
            if (top[i - 1] == sc) {
                diag = wadd(prev[i - 1], m);
                code |= NWO_MATCH;
            } else {
                diag = wsub(prev[i - 1], k);
            }
            const int32_t s = max3(up, left, diag);
            cur[i] = s;
            uint64_t c = 0;
            int arrows = 0;
            if (s == diag) { code |= NWO_DIAG; c += cprev[i - 1]; arrows++; }
            if (s == up)   { code |= NWO_UP;   c += cprev[i];     arrows++; }
            if (s == left) { code |= NWO_LEFT; c += ccur[i - 1];  arrows++; }
            ccur[i] = c;
            if (arrows > 1) branches++;
            /* score_cell_column(), needleman-wunsch.c:538-541 (tflag rule) */
            const int32_t a = iabs32(s);
            if (a > gabs) gabs = a;

            if (scores) scores[j * M + i] = s;
            if (codes) codes[j * M + i] = code;
            if (counts) counts[j * M + i] = c;
            if (prow) prow[(i - 1) >> 1] |= (uint8_t)((code & 7u) << (((i - 1) & 1) * 4));
            th = fnv_word(th, (uint32_t)s);
            th = fnv_word(th, code);
            ah = fnv_word(ah, code & 7u);
            ch = fnv_u64(ch, c);
            if (j == N - 1) lrh = fnv_u64(lrh, c);
            if (j == N - 1) lrdig += nwo_mix64((uint64_t)i, c);
            dword |= (uint32_t)(code & 7u) << (((i - 1) & 7) * 4);
            if (((i - 1) & 7) == 7 || i == M - 1) {
                adig += nwo_mix64(((uint64_t)j << 32) | (uint64_t)((i - 1) >> 3), dword);
                dword = 0;
            }
        }
        lch = fnv_u64(lch, ccur[M - 1]);
        if (M > 1) lcdig += nwo_mix64((uint64_t)j, ccur[M - 1]);

        int32_t *t = prev; prev = cur; cur = t;
        uint64_t *tc = cprev; cprev = ccur; ccur = tc;
    }

    if (res) {
        memset(res, 0, sizeof(*res));
        res->final_score = prev[M - 1];
        res->branch_count = branches;
        res->greatest_abs = gabs;
        res->table_hash = th;
        res->arrow_hash = ah;
        res->count = cprev[M - 1];
        res->count_hash = ch;
        res->lastrow_count_hash = lrh;
        res->lastcol_count_hash = lch;
        res->arrow_digest = adig;
        res->lastrow_count_digest = lrdig;
        res->lastcol_count_digest = lcdig;
    }
    free(prev); free(cur); free(cprev); free(ccur);
    return 0;
}

/* SURVEY.md 8d generator */
void nwo_generate(uint64_t seed, const char *alphabet, int alen, char *out, size_t n)
{
    uint64_t s = seed;
    for (size_t i = 0; i < n; i++) {
        s += 0x9E3779B97F4A7C15ULL;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        out[i] = alphabet[z % (uint64_t)alen];
    }
}

/*
 * Enumeration in the reference's order.  The reference keeps per-cell
 * *_done / src_direction state (needleman-wunsch.c:231-327) that is only live
 * on the current DFS path; an explicit stack of (i, j, next-direction) is the
 * same traversal: at each cell try diag, then left, then up.
 */
uint64_t nwo_enumerate(const char *top, int A, const char *side, int B,
                       const uint8_t *codes, uint64_t limit,
                       nwo_align_cb cb, void *user)
{
    const size_t M = (size_t)A + 1;
    const int maxlen = A + B + 1;
    char *X = (char *)malloc((size_t)maxlen);
    char *Y = (char *)malloc((size_t)maxlen);
    char *Xo = (char *)malloc((size_t)maxlen);
    char *Yo = (char *)malloc((size_t)maxlen);
    /* stack entry: which direction to try next at depth n (0 diag,1 left,2 up,3 done) */
    uint8_t *next = (uint8_t *)calloc((size_t)maxlen + 1, 1);
    uint8_t *came = (uint8_t *)calloc((size_t)maxlen + 1, 1);
    uint64_t found = 0;
    if (!X || !Y || !Xo || !Yo || !next || !came) goto out;

    int i = A, j = B, n = 0;
    next[0] = 0;
    for (;;) {
        if (i == 0 && j == 0) {
            /* X/Y were filled from the end of the alignment backwards */
            for (int t = 0; t < n; t++) { Xo[t] = X[n - 1 - t]; Yo[t] = Y[n - 1 - t]; }
            if (cb) cb(Xo, Yo, n, user);
            found++;
            if (limit && found >= limit) break;
        }
        const uint8_t code = codes[(size_t)j * M + (size_t)i];
        int moved = 0;
        while (next[n] < 3 && !moved) {
            const int dir = next[n]++;
            if (dir == 0 && (code & NWO_DIAG)) {
                X[n] = top[i - 1]; Y[n] = side[j - 1]; i--; j--; came[n + 1] = 0; moved = 1;
            } else if (dir == 1 && (code & NWO_LEFT)) {
                X[n] = top[i - 1]; Y[n] = '-'; i--; came[n + 1] = 1; moved = 1;
            } else if (dir == 2 && (code & NWO_UP)) {
                X[n] = '-'; Y[n] = side[j - 1]; j--; came[n + 1] = 2; moved = 1;
            }
        }
        if (moved) {
            n++;
            next[n] = 0;
            continue;
        }
        /* all directions done here: go back to the source cell */
        if (n == 0) break;
        switch (came[n]) {
        case 0: i++; j++; break;
        case 1: i++; break;
        default: j++; break;
        }
        n--;
    }
out:
    free(X); free(Y); free(Xo); free(Yo); free(next); free(came);
    return found;
}
