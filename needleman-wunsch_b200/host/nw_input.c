/*
 * nw_input.c -- sequence reader of the host shell.  Behaviour follows the
 * reference's read-sequences.c:43-122: characters up to the first isspace()
 * form a sequence; the buffer grows in 4096-byte steps; a stream error is
 * "fgetc failed"; running into EOF where a sequence must still follow is
 * "got EOF too early when reading input strings".
 */
#include "nw_input.h"

#include <ctype.h>

#include "nw_err.h"

#define NW_CHUNK 4096

static void stream_ok(FILE *in, int eof_allowed)
{
    nw_require(ferror(in) == 0, "fgetc failed");
    if (!eof_allowed) nw_require(feof(in) == 0, "got EOF too early when reading input strings");
}

static char *read_token(FILE *in, int eof_allowed)
{
    size_t cap = NW_CHUNK, len = 0;
    char *buf = (char *)malloc(cap);
    nw_require(buf != NULL, "malloc failed");
    for (;;) {
        const int ch = fgetc(in);
        if (ch == EOF || isspace(ch)) break;
        buf[len++] = (char)ch;
        if (len == cap) {
            cap += NW_CHUNK;
            buf = (char *)realloc(buf, cap);
            nw_require(buf != NULL, "realloc failed");
        }
    }
    stream_ok(in, eof_allowed);
    buf[len] = '\0';
    return buf;
}

void nw_read_pair(FILE *in, char **first, char **second)
{
    char *a = read_token(in, 0);
    int ch = ' ';
    while (isspace(ch)) ch = fgetc(in); /* skip the separator run */
    stream_ok(in, 0);
    nw_require(ungetc(ch, in) != EOF, "ungetc failed");
    char *b = read_token(in, 1);
    *first = a;
    *second = b;
}
