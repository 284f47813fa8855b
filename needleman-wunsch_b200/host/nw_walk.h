#ifndef NW_WALK_H
#define NW_WALK_H
#include <stdint.h>
#include "nwb.h"

/* Enumerate every optimal alignment in the reference's order (diag, then left,
 * then up; needleman-wunsch.c:209-331) and print each one the way
 * print_aligned_strings_and_counts() does (needleman-wunsch.c:137-182). */
void nw_enumerate_and_print(const nwb_table *t, const char *top, const char *side,
                            int quiet, int list_counts);

/* Cells on at least one optimal path (walk_table_cell_t.in_optimal_path under
 * tflag, needleman-wunsch.c:239-241): reachability over the arrows from the
 * bottom-right cell -- every arrow path ends in (0,0), so "reachable" equals
 * "visited by the reference's enumeration", in linear instead of exponential
 * time.  Returns a (A+1)*(B+1) byte map indexed [j*(A+1)+i]; caller frees. */
uint8_t *nw_mark_optimal_paths(const nwb_table *t);
#endif
