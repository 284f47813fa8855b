/*
 * nw_batch_cli.c -- batch front-end of the B200 score-table fill:
 *
 *   needleman-wunsch-batch [-h][-s] [-f sequence-file] m k d
 *
 * The reference reads exactly one pair (read-sequences.c:102-122) and is looped per
 * pair by its users: `needleman-wunsch -q -s m k d < pair`.  This program reads ANY
 * number of whitespace-separated strings -- s1 s2 s1 s2 ... -- (a sequence ends at the
 * first isspace(), as in the reference; runs of whitespace separate, so a batch cannot
 * hold the reference's empty first string; an odd number of strings, or none, is
 * "got EOF too early when reading input strings"), hands all pairs to ONE
 * nwb_fill_batch() call (include/nwb.h section 3; pairs are independent, a warp sweeps
 * two of them at a time) and prints, per pair and in input order, what that loop
 * prints on stderr:
 *     "%d optimal alignment%s\n"   (plural iff > 1; the low 32 bits of the count, as the
 *                                  reference's unsigned solution_count shown with %d)
 *     "Optimal score is %-d\n"
 * -- the first line only with -s (computation.c:271-281).  Always quiet (-q): alignments
 * of a batch are not enumerated.  A separate program so that the drop-in CLI keeps the
 * reference's exact option set.
 */
#include <ctype.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "nw_err.h"
#include "nwb.h"

const char *nw_prog = "needleman-wunsch-batch";

static void usage_and_exit(void)
{
    fputs("usage: needleman-wunsch-batch [-h][-s] [-f sequence-file] m k d\n"
          "Score many pairs of sequences (s1 s2 s1 s2 ...) with the Needleman-Wunsch algorithm\n"
          "operands:\n"
          "   m   match bonus\n"
          "   k   mismatch penalty\n"
          "   d   indel (gap) penalty\n"
          "options:\n"
          "  -f sequence-file\n"
          "       read the input strings from 'sequence-file' instead of standard input\n"
          "  -h   print this usage message\n"
          "  -s   also print the number of optimal alignments of every pair\n",
          stderr);
    exit(1);
}

struct bytes {
    char *p;
    size_t len, cap;
};

static void push(struct bytes *b, char c)
{
    if (b->len == b->cap) {
        b->cap = b->cap ? 2 * b->cap : 1 << 16;
        b->p = (char *)realloc(b->p, b->cap);
        nw_require(b->p != NULL, "realloc failed");
    }
    b->p[b->len++] = c;
}

int main(int argc, char **argv)
{
    int summary = 0;
    const char *path = NULL;
    nw_prog = argv[0];
    if (nw_prog[0] == '.' && nw_prog[1] == '/') nw_prog += 2;
    errno = 0;
    int opt;
    while ((opt = getopt(argc, argv, "f:hs")) != -1) {
        switch (opt) {
        case 'f': path = optarg; break;
        case 's': summary = 1; break;
        case 'h':
        default: usage_and_exit();
        }
    }
    const int operands = argc - optind;
    if (operands != 3) {
        nw_error("expected %d operands but received%s %d", 3, (operands > 3 || operands == 0) ? "" : " only", operands);
        usage_and_exit();
    }
    FILE *in = stdin;
    if (path) {
        in = fopen(path, "r");
        nw_require(in != NULL, "failed to open %s", path);
    }
    const int m = atoi(argv[optind]), k = atoi(argv[optind + 1]), d = atoi(argv[optind + 2]);

    /* tokens alternate top, side; tops and sides are concatenated with offset arrays (include/nwb.h) */
    struct bytes cat[2] = {{NULL, 0, 0}, {NULL, 0, 0}};
    int64_t *off[2] = {NULL, NULL};
    size_t n[2] = {0, 0}, cap = 0;
    int which = 0, in_token = 0;
    for (;;) {
        const int ch = fgetc(in);
        if (ch != EOF && !isspace(ch)) {
            if (!in_token) {
                if (n[0] + 2 > cap) {
                    cap = cap ? 2 * cap : 1024;
                    for (int w = 0; w < 2; w++) {
                        off[w] = (int64_t *)realloc(off[w], (cap + 1) * sizeof(int64_t));
                        nw_require(off[w] != NULL, "realloc failed");
                    }
                }
                off[which][n[which]] = (int64_t)cat[which].len;
                in_token = 1;
            }
            push(&cat[which], (char)ch);
            continue;
        }
        if (in_token) {
            n[which]++;
            which ^= 1;
            in_token = 0;
        }
        if (ch == EOF) break;
    }
    nw_require(ferror(in) == 0, "fgetc failed");
    nw_require(n[0] > 0 && n[0] == n[1], "got EOF too early when reading input strings");
    off[0][n[0]] = (int64_t)cat[0].len;
    off[1][n[1]] = (int64_t)cat[1].len;

    nwb_batch *b = NULL;
    const int rc = nwb_fill_batch(cat[0].p, off[0], cat[1].p, off[1], (int64_t)n[0], m, k, d,
                                  summary ? NWB_WANT_COUNT : 0u, 0, &b);
    if (rc != NWB_OK) {
        errno = 0;
        nw_require(0, "score-table fill failed: %s (%s)", nwb_strerror(rc), nwb_last_cuda_error());
    }
    for (size_t p = 0; p < n[0]; p++) {
        if (summary) {
            const unsigned count = (unsigned)nwb_batch_count_u64(b, (int64_t)p);
            fprintf(stderr, "%d optimal alignment%s\n", count, count > 1 ? "s" : "");
        }
        fprintf(stderr, "Optimal score is %-d\n", (int)nwb_batch_opt_score(b, (int64_t)p));
    }
    nwb_batch_free(b);
    free(cat[0].p); free(cat[1].p); free(off[0]); free(off[1]);
    return 0;
}
