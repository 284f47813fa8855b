/*
 * nw_err.h -- diagnostics of the host shell, reproducing the reference's
 * message format and exit behaviour (reference dbg.h:61-100, NDEBUG build):
 *     "<prog>: error: <message>: <strerror(errno)>\n"   then exit(1)
 * where a zero errno prints "\b\b \b" in place of the errno text (dbg.h:63).
 */
#ifndef NW_ERR_H
#define NW_ERR_H

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

extern const char *nw_prog;

#define NW_ERRNO_TEXT() (errno == 0 ? "\b\b \b" : strerror(errno))

#define nw_error(...)                                   \
    do {                                                \
        fprintf(stderr, "%s: error: ", nw_prog);        \
        fprintf(stderr, __VA_ARGS__);                   \
        fprintf(stderr, ": %s\n", NW_ERRNO_TEXT());     \
    } while (0)

#define nw_require(cond, ...)                           \
    do {                                                \
        if (!(cond)) {                                  \
            nw_error(__VA_ARGS__);                      \
            exit(1);                                    \
        }                                               \
    } while (0)

#endif
