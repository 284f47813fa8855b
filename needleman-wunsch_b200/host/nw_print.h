#ifndef NW_PRINT_H
#define NW_PRINT_H
#include <stdint.h>
#include "nwb.h"

extern int nw_color; /* -c */

/* One aligned character, coloured by its relation to its partner
 * (reference needleman-wunsch.c:101-119, format.h:72-100). */
void nw_put_aligned_char(char c, char partner);

/* The -t table (reference print-table.c:50-208): arrow row + score row per table
 * row, column width from the INTERIOR |score| maximum, ASCII or unicode arrows,
 * cells on an optimal path highlighted under -c. */
void nw_print_table(const nwb_table *t, const char *top, const char *side, const uint8_t *on_path, int unicode);
#endif
