/*
 * nw_print.c -- text output of the host shell: coloured alignment characters and
 * the -t score/arrow table, reading the GPU-produced table through the C ABI
 * accessors (nwb_score, nwb_arrows, nwb_greatest_abs_interior).  Output is
 * byte-identical to the reference's print-table.c / format.c.
 */
#include "nw_print.h"

#include <stdio.h>

int nw_color = 0;

/* SGR sequences of reference format.h:72-108 */
#define SGR_BOLD "\x1b[1m"
#define SGR_ON_PATH "\x1b[32;1m"
#define SGR_MATCH_ARROW "\x1b[36;1m"
#define SGR_MISMATCH "\x1b[31;1m"
#define SGR_GAP_ARROW "\x1b[33;1m"
#define SGR_RESET "\x1b[0m"

static void sgr(const char *seq)
{
    if (nw_color) fputs(seq, stdout);
}

void nw_put_aligned_char(char c, char partner)
{
    /* match and gap characters carry an empty format, mismatches are red/bold */
    if (c != partner && c != '-' && partner != '-') sgr(SGR_MISMATCH);
    else sgr("");
    printf("%c", c);
    sgr(SGR_RESET);
}

static int digits_plus_sign(int x)
{
    int w = 0;
    do {
        x /= 10;
        w++;
    } while (x != 0);
    return w + 1;
}

void nw_print_table(const nwb_table *t, const char *top, const char *side, const uint8_t *on_path, int unicode)
{
    const int A = nwb_top_len(t), B = nwb_side_len(t);
    const size_t W = (size_t)A + 1;
    const int cw = digits_plus_sign(nwb_greatest_abs_interior(t));
    const char *arrow_left = unicode ? "←" : "<";
    const char *arrow_up = unicode ? "↑" : "^";
    const char *arrow_diag = unicode ? "↖" : "\\";

    /* header: the top string (its bold is never reset, as in the reference) */
    sgr(SGR_BOLD);
    printf("*    %*s", cw, "-");
    for (int i = 0; i < A; i++) printf("    %*s%c", cw - 1, "", top[i]);
    printf("\n");

    for (int j = 0; j <= B; j++) {
        /* arrow row: diagonal and up arrows of every cell */
        printf(" ");
        for (int i = 0; i <= A; i++) {
            const unsigned code = nwb_arrows(t, i, j);
            const int hot = on_path[(size_t)j * W + (size_t)i];
            if (code & NWB_DIAG) {
                if (hot) sgr(top[i - 1] == side[j - 1] ? SGR_MATCH_ARROW : SGR_MISMATCH);
                printf("  %s ", arrow_diag);
                if (hot) sgr(SGR_RESET);
            } else {
                printf("    ");
            }
            if (code & NWB_UP) {
                if (hot) sgr(SGR_GAP_ARROW);
                printf("%*s", unicode ? cw + 2 : cw, arrow_up);
                if (hot) sgr(SGR_RESET);
            } else {
                printf("%*s", cw, "");
            }
        }
        printf("\n");
        /* score row: side character, then left arrow + score of every cell */
        sgr(SGR_BOLD);
        printf("%c", j == 0 ? '-' : side[j - 1]);
        sgr(SGR_RESET);
        for (int i = 0; i <= A; i++) {
            const unsigned code = nwb_arrows(t, i, j);
            const int hot = on_path[(size_t)j * W + (size_t)i];
            if (code & NWB_LEFT) {
                if (hot) sgr(SGR_GAP_ARROW);
                printf("  %s ", arrow_left);
                if (hot) sgr(SGR_RESET);
            } else {
                printf("    ");
            }
            if (hot) sgr(SGR_ON_PATH);
            printf("%+*d", cw, (int)nwb_score(t, i, j));
            if (hot) sgr(SGR_RESET);
        }
        printf("\n");
    }
}
