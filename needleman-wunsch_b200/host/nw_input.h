#ifndef NW_INPUT_H
#define NW_INPUT_H
#include <stdio.h>
/* Read two whitespace-separated sequences (reference read-sequences.c:102-122):
 * the first may be empty (leading whitespace), EOF before the second one starts
 * is fatal, anything after the second one is ignored. */
void nw_read_pair(FILE *in, char **first, char **second);
#endif
