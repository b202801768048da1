/* freesasa_b200/csrc/host_internal.h — shared by the C files of libfreesasa_b200_host.so (not installed). */
#ifndef FSB_HOST_INTERNAL_H
#define FSB_HOST_INTERNAL_H

#include "freesasa_b200_host.h"

/* Error/warning sink with the reference's conventions (src/util.c:36-102): "freesasa: warning: ..." and
 * "freesasa:<file>:<line>: error: ..." on the stream chosen with freesasa_set_err_out(), filtered by the
 * verbosity.  Returns `code` so that `return FAIL_MSG(...)` reads like the reference's `return fail_msg(...)`. */
int fsb_report(int code, const char *where, int line, const char *fmt, ...)
#if defined(__GNUC__)
    __attribute__((format(printf, 4, 5)))
#endif
    ;
#define FAIL_MSG(...) fsb_report(FREESASA_FAIL, __FILE__, __LINE__, __VA_ARGS__)
#define WARN_MSG(...) fsb_report(FREESASA_WARN, NULL, 0, __VA_ARGS__)
#define MEM_FAIL() FAIL_MSG("Out of memory")

/* first whitespace-delimited token of `s` (what the reference's sscanf(key, "%s", ...) extracts,
 * src/classifier.c:126-160): *begin..*begin+len, len 0 if `s` is all whitespace */
static inline int fsb_token(const char *s, const char **begin)
{
    const char *p = s, *q;
    while (*p == ' ' || (*p >= '\t' && *p <= '\r')) ++p;
    q = p;
    while (*q && !(*q == ' ' || (*q >= '\t' && *q <= '\r'))) ++q;
    *begin = p;
    return (int)(q - p);
}

#endif
