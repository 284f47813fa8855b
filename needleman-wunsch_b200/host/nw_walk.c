/*
 * nw_walk.c -- consumers of the GPU-produced arrow table: alignment enumeration
 * and optimal-path marking.  The reference keeps per-cell mutable walk state
 * (*_done, src_direction; walk-table.h:52-56) that is only live along the
 * current DFS path; here the same traversal runs on an explicit stack of
 * (direction to try next, direction we came by) per path position, reading the
 * 4-bit codes through nwb_arrows().
 */
#include "nw_walk.h"

#include <stdio.h>
#include <stdlib.h>

#include "nw_err.h"
#include "nw_print.h"

enum { TRY_DIAG = 0, TRY_LEFT = 1, TRY_UP = 2, TRY_NONE = 3 };

/* X and Y hold the alignment from its END backwards (position 0 = last column),
 * exactly like the reference's buffers; printing walks them from n-1 down to 0. */
static void emit_alignment(const char *X, const char *Y, int n, int quiet, int list_counts)
{
    int same = 0, differ = 0, gaps = 0;
    for (int p = n - 1; p >= 0; p--) {
        if (!quiet) nw_put_aligned_char(X[p], Y[p]);
        if (list_counts) {
            if (X[p] == Y[p]) same++;
            else if (X[p] == '-' || Y[p] == '-') gaps++;
            else differ++;
        }
    }
    if (!quiet) {
        printf("\n");
        for (int p = n - 1; p >= 0; p--) nw_put_aligned_char(Y[p], X[p]);
        printf("\n");
    }
    if (list_counts)
        printf("%d match%s, %d mismatch%s, %d indel%s\n", same, same == 1 ? "" : "es", differ,
               differ == 1 ? "" : "es", gaps, gaps == 1 ? "" : "s");
    printf("\n");
}

void nw_enumerate_and_print(const nwb_table *t, const char *top, const char *side, int quiet, int list_counts)
{
    const int A = nwb_top_len(t), B = nwb_side_len(t);
    const size_t cap = (size_t)A + (size_t)B + 2;
    char *X = (char *)malloc(cap), *Y = (char *)malloc(cap);
    unsigned char *next_dir = (unsigned char *)calloc(cap, 1), *came_by = (unsigned char *)calloc(cap, 1);
    nw_require(X && Y && next_dir && came_by, "malloc failed");

    int i = A, j = B, n = 0;
    const int print_any = (!quiet) || list_counts;
    for (;;) {
        if (i == 0 && j == 0 && print_any) emit_alignment(X, Y, n, quiet, list_counts);
        const unsigned code = nwb_arrows(t, i, j);
        int stepped = 0;
        while (!stepped && next_dir[n] != TRY_NONE) {
            const int dir = next_dir[n]++;
            if (dir == TRY_DIAG && (code & NWB_DIAG)) {
                X[n] = top[i - 1]; Y[n] = side[j - 1]; i--; j--; stepped = 1;
            } else if (dir == TRY_LEFT && (code & NWB_LEFT)) {
                X[n] = top[i - 1]; Y[n] = '-'; i--; stepped = 1;
            } else if (dir == TRY_UP && (code & NWB_UP)) {
                X[n] = '-'; Y[n] = side[j - 1]; j--; stepped = 1;
            }
            if (stepped) came_by[n + 1] = (unsigned char)dir;
        }
        if (stepped) {
            n++;
            next_dir[n] = TRY_DIAG;
            continue;
        }
        if (n == 0) break; /* every direction of the start cell is exhausted */
        if (came_by[n] == TRY_DIAG) { i++; j++; }
        else if (came_by[n] == TRY_LEFT) i++;
        else j++;
        n--;
    }
    free(X); free(Y); free(next_dir); free(came_by);
}

uint8_t *nw_mark_optimal_paths(const nwb_table *t)
{
    const int A = nwb_top_len(t), B = nwb_side_len(t);
    const size_t W = (size_t)A + 1;
    uint8_t *mark = (uint8_t *)calloc(W * ((size_t)B + 1), 1);
    nw_require(mark != NULL, "malloc failed");
    mark[(size_t)B * W + (size_t)A] = 1;
    /* arrows only point up/left/diag, so one sweep from the bottom-right suffices */
    for (int j = B; j >= 0; j--) {
        for (int i = A; i >= 0; i--) {
            if (!mark[(size_t)j * W + (size_t)i]) continue;
            const unsigned code = nwb_arrows(t, i, j);
            if (code & NWB_DIAG) mark[(size_t)(j - 1) * W + (size_t)(i - 1)] = 1;
            if (code & NWB_LEFT) mark[(size_t)j * W + (size_t)(i - 1)] = 1;
            if (code & NWB_UP) mark[(size_t)(j - 1) * W + (size_t)i] = 1;
        }
    }
    return mark;
}
