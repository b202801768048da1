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
/* per-thread message capture (see fsb_report): while fsb_capture_current is set, messages are appended to it */
struct fsb_capture {
    char *text;
    size_t len, cap;
};
extern __thread struct fsb_capture *fsb_capture_current;
void fsb_capture_flush(struct fsb_capture *c); /* writes the captured text to the error stream and empties the buffer */
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

/* workers.c */
int fsb_hardware_threads(void);
void fsb_parallel_run(int n_parts, void (*fn)(int part, int n_parts, void *arg), void *arg);

/* ---- structure storage (ingest.c), shared with the result tree (areas.c) --------------------------------------- */
#define LINE_MAX_STRL 120 /* PDB_MAX_LINE_STRL, src/pdb.h:22: fgets() buffer of the reference, longer lines are split */

/* the bytes a set of structures was read from, shared by reference count (freesasa_structure_array() makes several
 * structures from one file) */
struct shared_text {
    char *data;
    long len;
    int refs;
};
void fsb_text_release(struct shared_text *t); /* drops one reference */

/* fixed-width, NUL-terminated label fields of one atom (widths: src/pdb.h:17-20, src/structure.c:30-32) */
struct atom_label {
    char name[5];
    char res_name[4];
    char res_number[6];
    char symbol[3];
    char chain[4];
};

struct freesasa_structure {
    int n, cap;
    coord_t coord; /* coord.xyz owned, 3*cap doubles; what freesasa_calc() receives */
    double *radius;
    struct atom_label *label;
    int *res_index;
    unsigned char *cls; /* bits 0-1: freesasa_atom_class, bit 2: backbone atom (freesasa_atom_is_backbone of the name) */
    /* PDB lines are not copied while reading: the structure keeps a reference to the text it was read from and a
     * slice per atom; NUL-terminated copies are made only if somebody asks for them (atom_pdb_line) */
    struct shared_text *text;
    long *line_at;           /* offset into text->data, -1 = the atom did not come from a PDB line */
    unsigned char *line_len; /* <= 119 */
    char *volatile lines;    /* lazily built: n strings of stride LINE_MAX_STRL */
    /* residues */
    int n_res, cap_res;
    int *res_first;
    freesasa_nodearea *res_ref;
    unsigned char *res_has_ref;
    /* chains, in order of first appearance */
    int n_chains, cap_chains;
    char (*chain_label)[4];
    char *short_labels;
    int *chain_first;
    char *classifier_name;
    const freesasa_classifier *last_classifier; /* classifier of the previous add: skips re-registering its name */
    /* direct-mapped memo of freesasa_atom_is_backbone() by the four name bytes (a structure has a few dozen names) */
    unsigned int bb_key[64];
    unsigned char bb_val[64];
    int model;
};


#define FSB_CLS_CLASS(c) ((c)&3)
#define FSB_CLS_BACKBONE(c) (((c) >> 2) & 1)

#endif
