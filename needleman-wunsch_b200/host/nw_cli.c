/*
 * nw_cli.c -- drop-in command line of skotchandsoda/needleman-wunsch on top of
 * the B200 score-table fill:
 *
 *   needleman-wunsch [-c][-h][-l][-q][-s][-t][-u] [-p num-threads] [-f sequence-file] m k d
 *
 * Same options, operands, stdout/stderr text and exit codes as the reference's
 * main()/needleman_wunsch() (needleman-wunsch.c:654-795).  The fill itself is the
 * single call nwb_fill() (include/nwb.h) -- there is no CPU fill in this program:
 * without a CUDA device it exits with an error.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "nw_err.h"
#include "nw_input.h"
#include "nw_print.h"
#include "nw_walk.h"
#include "nwb.h"

const char *nw_prog = "needleman-wunsch";

static void usage_and_exit(void)
{
    fputs("usage: needleman-wunsch [-c][-h][-l][-q][-s][-t][-u]\n"
          "                        [-p num-threads] [-f sequence-file] m k d\n"
          "Align two sequences with the Needleman-Wunsch algorithm\n"
          "operands:\n"
          "   m   match bonus\n"
          "   k   mismatch penalty\n"
          "   d   indel (gap) penalty\n"
          "options:\n"
          "  -c   color the output with ANSI escape sequences\n"
          "  -f sequence-file\n"
          "       read the input strings from 'sequence-file' instead of standard input\n"
          "  -h   print this usage message\n"
          "  -l   list match, mismatch, and indel counts for each alignment pair\n"
          "  -p num-threads\n"
          "       parallelize the computation with 'num-threads' threads (must be >1)\n"
          "  -q   be quiet and don't print the aligned strings\n"
          "  -s   summarize the algorithm's run\n"
          "  -t   print the scores table; only useful for shorter input strings\n"
          "  -u   use unicode arrows when printing the scores table\n",
          stderr);
    exit(1);
}

int main(int argc, char **argv)
{
    int list_counts = 0, quiet = 0, summary = 0, table = 0, unicode = 0, workers = 1;
    const char *path = NULL;

    /* reference dbg.c:7-15: a leading "./" is dropped from the program name */
    nw_prog = argv[0];
    if (nw_prog[0] == '.' && nw_prog[1] == '/') nw_prog += 2;
    errno = 0;

    int opt;
    while ((opt = getopt(argc, argv, "cf:hlp:qstu")) != -1) {
        switch (opt) {
        case 'c': nw_color = 1; break;
        case 'f': path = optarg; break;
        case 'l': list_counts = 1; break;
        case 'p': {
            /* the reference's worker count (needleman-wunsch.c:738-742, same check and message); here the
             * workers are GPUs: the table is split into column strips over up to that many devices */
            const int threads = atoi(optarg);
            nw_require(threads > 1, "num-threads == %d; num-threads must be greater than 1", threads);
            workers = threads;
            break;
        }
        case 'q': quiet = 1; break;
        case 's': summary = 1; break;
        case 't': table = 1; break;
        case 'u': unicode = 1; break;
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
    char *top, *side;
    nw_read_pair(in, &top, &side);
    const int m = atoi(argv[optind]), k = atoi(argv[optind + 1]), d = atoi(argv[optind + 2]);

    /* what the consumers below will read decides what the fill must deliver */
    unsigned flags = 0;
    if (summary) flags |= NWB_WANT_COUNT;
    if (table) flags |= NWB_WANT_SCORES | NWB_TRACK_ABS | NWB_WANT_ARROWS_HOST;
    if (!quiet || list_counts) flags |= NWB_WANT_ARROWS_HOST;
    nwb_table *t = NULL;
    /* -p N: column strips over min(N, devices present) GPUs, boundary columns pipelined over NVLink peer
     * memory (include/nwb.h, nwb_fill_on).  The int32 score matrix behind -t lives on one device, and a
     * box whose GPUs cannot map one another's memory still has every single GPU: both fill on device 0. */
    int gpus = 1;
    if (workers > 1 && !table) {
        const int ndev = nwb_device_count();
        gpus = workers < ndev ? workers : ndev;
        if (gpus < 1) gpus = 1;
    }
    int rc = NWB_ERR_CUDA;
    if (gpus > 1) rc = nwb_fill_on(top, (int)strlen(top), side, (int)strlen(side), m, k, d, flags, 0, gpus, &t);
    if (rc != NWB_OK) rc = nwb_fill(top, (int)strlen(top), side, (int)strlen(side), m, k, d, flags, &t);
    if (rc != NWB_OK) {
        errno = 0;
        nw_require(0, "score-table fill failed: %s (%s)", nwb_strerror(rc), nwb_last_cuda_error());
    }

    /* reference needleman-wunsch.c:667-669: the walk prints the alignments; its
     * count and path marks come from the fused DP / a reachability sweep here */
    if (!quiet || list_counts) nw_enumerate_and_print(t, top, side, quiet, list_counts);
    if (summary) {
        const unsigned count = (unsigned)nwb_count_u64(t); /* unsigned int solution_count, printed with %d */
        fprintf(stderr, "%d optimal alignment%s\n", count, count > 1 ? "s" : "");
        fprintf(stderr, "Optimal score is %-d\n", (int)nwb_opt_score(t));
    }
    if (table) {
        if (!quiet || summary || list_counts) printf("\n");
        uint8_t *on_path = nw_mark_optimal_paths(t);
        nw_print_table(t, top, side, on_path, unicode);
        free(on_path);
    }
    nwb_free(t);
    free(top);
    free(side);
    return 0;
}
