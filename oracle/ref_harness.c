/*
 * ref_harness.c -- links the UNMODIFIED reference objects (compiled where
 * they lie under /root/reference by oracle/Makefile, outputs only into
 * oracle/_ref/) and exposes what compute_table_scores() left in the
 * reference's own tables.  TEST INFRASTRUCTURE ONLY (see nw_oracle.h).
 *
 * Nothing here restates the algorithm: every number returned is read out of
 * the reference's score_table_t / walk_table_t after the reference's own
 *   alloc_computation()/init_computation()   computation.c:51,145
 *   compute_table_scores()                   needleman-wunsch.c:583
 *   construct_alignments()                   needleman-wunsch.c:356
 * ran.  needleman-wunsch.c is compiled with -Dmain=nw_ref_main so that its
 * objects link into this library next to the reference CLI binary.
 */
#include <stdint.h>
#include <string.h>
#include <time.h>

#include "computation.h" /* from -I/root/reference */

#include "nw_oracle.h"

/* defined in the reference's needleman-wunsch.h / needleman-wunsch.c */
extern int lflag, qflag, sflag, tflag, uflag;
extern char *prog; /* dbg.h:44 */
extern void compute_table_scores(computation_t *C);
extern void construct_alignments(computation_t *C);

#define FNV_OFFSET 0xcbf29ce484222325ULL
#define FNV_PRIME 0x100000001b3ULL

static inline uint64_t fnv_word(uint64_t h, uint32_t w)
{
    h ^= (uint64_t)w;
    h *= FNV_PRIME;
    return h;
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/*
 * Run the reference fill on (top, side).  threads = the reference's -p value
 * (1 = default serial path).  want_tflag mirrors `-t` (enables
 * greatest_abs_val tracking, needleman-wunsch.c:538-541).
 * scores/codes (nullable) receive the full (B+1)x(A+1) row-major tables in the
 * nw_oracle.h convention.  enumerate != 0 additionally runs the reference's
 * alignment enumeration quietly and returns its solution_count in res->count
 * (exponential: small inputs only); otherwise res->count = 0.
 * fill_seconds / total_seconds (nullable): wall time of compute_table_scores()
 * alone / of alloc+init+fill+free.
 */
int nwref_fill(const char *top, const char *side, int m, int k, int d,
               int threads, int want_tflag, int enumerate,
               int32_t *scores, uint8_t *codes,
               nwo_result *res, double *fill_seconds, double *total_seconds)
{
    static char progname[] = "nwref";
    prog = progname;
    lflag = 0; sflag = 0; uflag = 0;
    qflag = 1;
    tflag = want_tflag ? 1 : 0;

    const double t0 = now_s();
    computation_t *C = alloc_computation();
    init_computation(C, (char *)top, (char *)side, m, k, d, (unsigned)threads);
    const double t1 = now_s();
    compute_table_scores(C);
    const double t2 = now_s();

    const int M = C->score_table->M, N = C->score_table->N;
    if (res) {
        memset(res, 0, sizeof(*res));
        uint64_t th = FNV_OFFSET, ah = FNV_OFFSET, adig = 0;
        for (int j = 0; j < N; j++) {
            uint32_t dword = 0;
            for (int i = 0; i < M; i++) {
                const score_table_cell_t *sc = &C->score_table->cells[i][j];
                const walk_table_cell_t *wc = &C->walk_table->cells[i][j];
                const uint32_t code = (uint32_t)(wc->diag | (wc->left << 1) | (wc->up << 2) | (sc->match << 3));
                th = fnv_word(th, (uint32_t)sc->score);
                th = fnv_word(th, code);
                ah = fnv_word(ah, code & 7u);
                if (scores) scores[(size_t)j * M + i] = sc->score;
                if (codes) codes[(size_t)j * M + i] = (uint8_t)code;
                /* nwo_result.arrow_digest over the reference's own walk table (interior cells) */
                if (i >= 1 && j >= 1) {
                    dword |= (code & 7u) << (((i - 1) & 7) * 4);
                    if (((i - 1) & 7) == 7 || i == M - 1) {
                        adig += nwo_mix64(((uint64_t)j << 32) | (uint64_t)((i - 1) >> 3), dword);
                        dword = 0;
                    }
                }
            }
        }
        res->final_score = C->score_table->cells[M - 1][N - 1].score;
        res->branch_count = get_branch_count(C->walk_table, (unsigned)threads);
        res->greatest_abs = C->score_table->greatest_abs_val;
        res->table_hash = th;
        res->arrow_hash = ah;
        res->arrow_digest = adig;
    }
    if (enumerate) {
        tflag = 0;
        construct_alignments(C);
        if (res) res->count = (uint64_t)get_solution_count(C);
    }
    const double t3 = now_s();
    free_computation(C);
    const double t4 = now_s();
    if (fill_seconds) *fill_seconds = t2 - t1;
    if (total_seconds) *total_seconds = (t2 - t0) + (t4 - t3);
    return 0;
}

/* sizeof probes (104 B + 32 B per cell) */
int nwref_sizeof_score_cell(void) { return (int)sizeof(score_table_cell_t); }
int nwref_sizeof_walk_cell(void) { return (int)sizeof(walk_table_cell_t); }
