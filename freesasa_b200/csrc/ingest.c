/* freesasa_b200/csrc/ingest.c — PDB text -> structure (coordinates, radii, labels) for the SASA engine.
 *
 * Scope row f-1 of SURVEY.md §8(f): once the calculation takes ~1 ms, reading the input dominates (the reference
 * needs 92 ms to read the 13 928 atoms of 2isk).  This file provides the reference's structure API
 * (src/freesasa.h:749-1446) with the same observable behaviour as src/structure.c + src/pdb.c, on a different
 * machine model:
 *
 *   reference                                            here
 *   ---------------------------------------------------  ------------------------------------------------------
 *   fgets() per line, ftell() per line                   the byte range is read with ONE fread; lines are slices
 *   malloc(struct atom) + strdup(line) per atom,         structure-of-arrays with geometric growth: xyz, radius,
 *   realloc of the pointer arrays every 512 atoms         fixed-width label records, one arena for the PDB lines
 *   3 classifier searches per atom, each malloc +         one hashed lookup per atom (radii.c)
 *   sscanf + linear strcmp scans
 *   sscanf("%lf%lf%lf") for the coordinates              exact decimal fast path (integer mantissa / power of ten,
 *                                                        one correctly rounded division == strtod), sscanf only for
 *                                                        exotic tokens (exponents, hex, inf/nan, > 15 digits)
 *   linear scan over all chains per atom                 previous atom's chain first
 *   one thread                                           files >= 4 MB fetched with pread slices, ranges >= 1 MB parsed as
 *                                                        one chunk per core and stitched; the models / chains of
 *                                                        freesasa_structure_array() parsed concurrently (messages replayed
 *                                                        in file order: the output is the serial one)
 *
 * Everything observable is kept, including the quirks: which lines count as atoms (src/structure.c:662-666), the
 * hydrogen test that only fires on lines long enough to carry an element symbol (src/pdb.c:260-283), "first
 * alternate location wins" (src/structure.c:672-679), residue = change of residue number or chain between
 * consecutive atoms (src/structure.c:484-496), chain = first appearance of a label (src/structure.c:459-478),
 * MODEL/ENDMDL handling, the file-range arithmetic of freesasa_structure_array(), messages and return codes.
 * The coordinates and radii are bit-identical to the reference's (tests/test_ingest.py).
 */
#include "host_internal.h"

#include <assert.h>
#include <errno.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>


int fsb_classifier_row(const freesasa_classifier *c, const char *res_name, const char *atom_name);
double fsb_classifier_row_radius(const freesasa_classifier *c, int row);
int fsb_classifier_row_class(const freesasa_classifier *c, int row);

void fsb_text_release(struct shared_text *t)
{
    if (t && __atomic_sub_fetch(&t->refs, 1, __ATOMIC_ACQ_REL) == 0) {
        free(t->data);
        free(t);
    }
}

/* ---- storage ---------------------------------------------------------------------------------------- */
static int reserve_atoms(freesasa_structure *s, int cap)
{
    void *p;
    if (cap <= s->cap) return FREESASA_SUCCESS;
#define GROW(field, bytes)                                \
    if (!(p = realloc(s->field, (size_t)cap * (bytes)))) return MEM_FAIL(); \
    s->field = p
    GROW(coord.xyz, 3 * sizeof(double));
    GROW(radius, sizeof(double));
    GROW(label, sizeof(struct atom_label));
    GROW(res_index, sizeof(int));
    GROW(cls, 1);
    GROW(line_at, sizeof(long));
    GROW(line_len, 1);
#undef GROW
    s->cap = cap;
    return FREESASA_SUCCESS;
}

freesasa_structure *freesasa_structure_new(void)
{
    freesasa_structure *s = calloc(1, sizeof *s);
    if (s == NULL) {
        MEM_FAIL();
        return NULL;
    }
    s->model = 1;
    s->coord.is_linked = 0;
    return s;
}

void freesasa_structure_free(freesasa_structure *s)
{
    if (s == NULL) return;
    free(s->coord.xyz);
    free(s->radius);
    free(s->label);
    free(s->res_index);
    free(s->cls);
    free(s->line_at);
    free(s->line_len);
    free(s->lines);
    fsb_text_release(s->text);
    free(s->res_first);
    free(s->res_ref);
    free(s->res_has_ref);
    free(s->chain_label);
    free(s->short_labels);
    free(s->chain_first);
    free(s->classifier_name);
    free(s);
}

/* ---- adding one atom (structure_add_atom(), src/structure.c:568-631) ------------------------------------------ */
static void copy_field(char *dst, size_t cap, const char *src)
{
    size_t n = strlen(src);
    if (n >= cap) n = cap - 1; /* snprintf(dst, cap, "%s", src) of the reference */
    memcpy(dst, src, n);
    dst[n] = '\0';
}

/* guess_symbol(), src/structure.c:419-445; characters past the end of a short name read as NUL */
static int guess_symbol(char symbol[3], const char *name)
{
    char n4[4] = {0, 0, 0, 0};
    int i;
    for (i = 0; i < 4 && name[i]; ++i) n4[i] = name[i];
    if (n4[0] == ' ' || (n4[0] >= '1' && n4[0] <= '9')) {
        symbol[0] = ' ';
        symbol[1] = n4[1];
        symbol[2] = '\0';
    } else if (n4[3] == ' ') {
        symbol[0] = n4[0];
        symbol[1] = n4[0] ? n4[1] : '\0';
        symbol[2] = '\0';
    } else {
        symbol[0] = ' ';
        symbol[1] = n4[0];
        symbol[2] = '\0';
        return WARN_MSG("guessing that atom '%s' is symbol '%s'", name, symbol);
    }
    return FREESASA_SUCCESS;
}

static int register_classifier(freesasa_structure *s, const freesasa_classifier *classifier)
{
    const char *name = freesasa_classifier_name(classifier);
    if (s->last_classifier == classifier && s->classifier_name) return FREESASA_SUCCESS;
    s->last_classifier = classifier;
    if (s->classifier_name == NULL) {
        if (!(s->classifier_name = strdup(name))) return MEM_FAIL();
    } else if (strcmp(s->classifier_name, name) != 0) {
        free(s->classifier_name);
        if (!(s->classifier_name = strdup(FREESASA_CONFLICTING_CLASSIFIERS))) return MEM_FAIL();
        return FREESASA_WARN;
    }
    return FREESASA_SUCCESS;
}

/* Returns FREESASA_SUCCESS (atom stored), FREESASA_WARN (atom skipped) or FREESASA_FAIL. */
#define ROW_UNSET (-2)
static int add_atom(freesasa_structure *s, const struct atom_label *a, const double xyz[3],
                    const freesasa_classifier *classifier, int options, const char *line, size_t line_len, int row)
{
    /* `a` is zero-padded (label fields compare with memcmp); `line` (optional) points into s->text; `row` is the
     * classifier row of the atom if the caller already knows it, else ROW_UNSET */
    int i;
    double r;

    if ((options & FREESASA_SKIP_UNKNOWN) && (options & FREESASA_HALT_AT_UNKNOWN)) options &= ~FREESASA_SKIP_UNKNOWN;
    if (classifier == NULL) classifier = &freesasa_default_classifier;
    if (s->last_classifier != classifier || !s->classifier_name) register_classifier(s, classifier);

    if (row == ROW_UNSET) row = fsb_classifier_row(classifier, a->res_name, a->name);
    if (options & FREESASA_RADIUS_FROM_OCCUPANCY) {
        r = 1; /* replaced by the caller */
    } else if (row >= 0) {
        r = fsb_classifier_row_radius(classifier, row);
    } else if (options & FREESASA_HALT_AT_UNKNOWN) {
        FAIL_MSG("atom '%s %s' unknown", a->res_name, a->name);
        return FAIL_MSG("halting at unknown atom");
    } else if (options & FREESASA_SKIP_UNKNOWN) {
        return WARN_MSG("skipping unknown atom '%s %s'", a->res_name, a->name);
    } else {
        r = freesasa_guess_radius(a->symbol);
        if (r < 0) {
            r = +0.;
            WARN_MSG("atom '%s %s' unknown and can't guess radius of symbol '%s', assigning radius 0 A", a->res_name,
                     a->name, a->symbol);
        } else {
            WARN_MSG("atom '%s %s' unknown, guessing element is '%s', and radius %.3f A", a->res_name, a->name, a->symbol, r);
        }
    }

    if (s->n == s->cap && reserve_atoms(s, s->cap ? 2 * s->cap : 1024)) return FAIL_MSG("%s", "");
    i = s->n;
    s->coord.xyz[3 * i] = xyz[0];
    s->coord.xyz[3 * i + 1] = xyz[1];
    s->coord.xyz[3 * i + 2] = xyz[2];

    /* new chain?  (structure_add_chain(), src/structure.c:459-478: a label is registered the first time it is seen) */
    if (i == 0 || memcmp(a->chain, s->label[i - 1].chain, sizeof a->chain) != 0) {
        int k, known = 0;
        for (k = 0; k < s->n_chains && !known; ++k) known = strncmp(s->chain_label[k], a->chain, 4) == 0;
        if (!known) {
            if (s->n_chains == s->cap_chains) {
                const int cap = s->cap_chains ? 2 * s->cap_chains : 64;
                void *p;
                if (!(p = realloc(s->chain_label, (size_t)cap * 4))) return MEM_FAIL();
                s->chain_label = p;
                if (!(p = realloc(s->short_labels, (size_t)cap + 1))) return MEM_FAIL();
                s->short_labels = p;
                if (!(p = realloc(s->chain_first, (size_t)cap * sizeof(int)))) return MEM_FAIL();
                s->chain_first = p;
                s->cap_chains = cap;
            }
            memcpy(s->chain_label[s->n_chains], a->chain, 4);
            s->short_labels[s->n_chains] = a->chain[0];
            s->short_labels[s->n_chains + 1] = '\0';
            s->chain_first[s->n_chains] = i;
            ++s->n_chains;
        }
    }
    /* new residue?  (structure_add_residue(), src/structure.c:480-516) */
    if (s->n_res == 0 || (i > 0 && (memcmp(a->res_number, s->label[i - 1].res_number, sizeof a->res_number) ||
                                    memcmp(a->chain, s->label[i - 1].chain, sizeof a->chain)))) {
        const freesasa_nodearea *ref;
        if (s->n_res == s->cap_res) {
            const int cap = s->cap_res ? 2 * s->cap_res : 256;
            void *p;
            if (!(p = realloc(s->res_first, (size_t)cap * sizeof(int)))) return MEM_FAIL();
            s->res_first = p;
            if (!(p = realloc(s->res_ref, (size_t)cap * sizeof(freesasa_nodearea)))) return MEM_FAIL();
            s->res_ref = p;
            if (!(p = realloc(s->res_has_ref, (size_t)cap))) return MEM_FAIL();
            s->res_has_ref = p;
            s->cap_res = cap;
        }
        s->res_first[s->n_res] = i;
        ref = freesasa_classifier_residue_reference(classifier, a->res_name);
        s->res_has_ref[s->n_res] = ref != NULL;
        if (ref) s->res_ref[s->n_res] = *ref;
        ++s->n_res;
    }
    s->label[i] = *a;
    {
        unsigned int key;
        unsigned slot;
        int bb;
        memcpy(&key, a->name, 4);
        slot = (key * 2654435761u) >> 26;
        if (s->bb_key[slot] == key && key != 0) {
            bb = s->bb_val[slot];
        } else {
            bb = freesasa_atom_is_backbone(a->name);
            s->bb_key[slot] = key;
            s->bb_val[slot] = (unsigned char)bb;
        }
        s->cls[i] = (unsigned char)((row >= 0 ? fsb_classifier_row_class(classifier, row) : FREESASA_ATOM_UNKNOWN) | (bb << 2));
    }
    s->res_index[i] = s->n_res - 1;
    s->radius[i] = r;
    s->line_at[i] = line ? line - s->text->data : -1;
    s->line_len[i] = (unsigned char)line_len;
    s->n = i + 1;
    s->coord.n = s->n;
    return FREESASA_SUCCESS;
}

/* structure_add_atom_wopt_impl(), src/structure.c:723-767 */
static int add_atom_wopt(freesasa_structure *s, const char *atom_name, const char *residue_name, const char *residue_number,
                         const char *symbol, const char *chain_label, double x, double y, double z,
                         const freesasa_classifier *classifier, int options)
{
    struct atom_label a;
    const double v[3] = {x, y, z};
    char my_symbol[3] = {0, 0, 0};
    int ret, warn = 0;

    assert(s);
    assert(atom_name);
    assert(residue_name);
    assert(residue_number);
    assert(chain_label);

    memset(&a, 0, sizeof a);
    options &= ~FREESASA_RADIUS_FROM_OCCUPANCY; /* cannot be used here */
    if (symbol != NULL)
        copy_field(my_symbol, sizeof my_symbol, symbol);
    else if (guess_symbol(my_symbol, atom_name) == FREESASA_WARN && (options & FREESASA_SKIP_UNKNOWN))
        ++warn;
    copy_field(a.name, sizeof a.name, atom_name);
    copy_field(a.res_name, sizeof a.res_name, residue_name);
    copy_field(a.res_number, sizeof a.res_number, residue_number);
    copy_field(a.symbol, sizeof a.symbol, my_symbol);
    copy_field(a.chain, sizeof a.chain, chain_label);
    ret = add_atom(s, &a, v, classifier, options, NULL, 0, ROW_UNSET);
    if (!ret && warn) return FREESASA_WARN;
    return ret;
}

int freesasa_structure_add_atom_wopt(freesasa_structure *structure, const char *atom_name, const char *residue_name,
                                     const char *residue_number, char chain_label, double x, double y, double z,
                                     const freesasa_classifier *classifier, int options)
{
    const char label[2] = {chain_label, '\0'};
    return add_atom_wopt(structure, atom_name, residue_name, residue_number, NULL, label, x, y, z, classifier, options);
}

int freesasa_structure_add_atom(freesasa_structure *structure, const char *atom_name, const char *residue_name,
                                const char *residue_number, char chain_label, double x, double y, double z)
{
    const char label[2] = {chain_label, '\0'};
    return add_atom_wopt(structure, atom_name, residue_name, residue_number, NULL, label, x, y, z, NULL, 0);
}

/* freesasa_structure_add_cif_atom() / _lcl(), src/structure.c:793-829: what an mmCIF reader (the reference's is the
 * gemmi-based src/cif.cc, outside this library) calls per _atom_site row */
static int add_cif_atom(freesasa_structure *structure, const char *atom_id, const char *comp_id, const char *seq_id,
                        const char *ins_code, const char *symbol, const char *chain, double x, double y, double z,
                        const freesasa_classifier *classifier, int options)
{
    char res_number[6];
    if (ins_code[0] != '?')
        snprintf(res_number, sizeof res_number, "%s%c", seq_id, ins_code[0]);
    else
        snprintf(res_number, sizeof res_number, "%s", seq_id);
    return add_atom_wopt(structure, atom_id, comp_id, res_number, symbol, chain, x, y, z, classifier, options);
}
int freesasa_structure_add_cif_atom(freesasa_structure *structure, freesasa_cif_atom *atom, const freesasa_classifier *classifier,
                                    int options)
{
    const char label[2] = {atom->auth_asym_id, '\0'};
    return add_cif_atom(structure, atom->auth_atom_id, atom->auth_comp_id, atom->auth_seq_id, atom->pdbx_PDB_ins_code,
                        atom->type_symbol, label, atom->Cartn_x, atom->Cartn_y, atom->Cartn_z, classifier, options);
}
int freesasa_structure_add_cif_atom_lcl(freesasa_structure *structure, freesasa_cif_atom_lcl *atom,
                                        const freesasa_classifier *classifier, int options)
{
    return add_cif_atom(structure, atom->auth_atom_id, atom->auth_comp_id, atom->auth_seq_id, atom->pdbx_PDB_ins_code,
                        atom->type_symbol, atom->auth_asym_id, atom->Cartn_x, atom->Cartn_y, atom->Cartn_z, classifier, options);
}

/* ---- PDB text ---------------------------------------------------------------------------------------- */
/* Large regular files are read in slices by several threads (pread: the copy out of the page cache and the first touch
 * of the destination pages are the cost, and both parallelise).  Returns the bytes read or -1 to ask for a plain fread. */
#define SLURP_PARALLEL_MIN_BYTES (4L << 20)
struct slice_job {
    int fd;
    char *dst;
    long offset, bytes, got;
};
static void *slice_read(void *arg)
{
    struct slice_job *j = arg;
    j->got = 0;
    while (j->got < j->bytes) {
        const ssize_t n = pread(j->fd, j->dst + j->got, (size_t)(j->bytes - j->got), (off_t)(j->offset + j->got));
        if (n <= 0) break;
        j->got += n;
    }
    return NULL;
}
static long slurp_parallel(FILE *f, char *dst, long len)
{
    struct slice_job job[16];
    pthread_t thread[16];
    const int fd = fileno(f);
    int n = fsb_hardware_threads(), k, started = 0;
    long total = 0;
    if (fd < 0 || n < 2) return -1;
    if (n > 16) n = 16;
    if (n > len / (1L << 20)) n = (int)(len / (1L << 20));
    if (n < 2) return -1;
    for (k = 0; k < n; ++k) {
        job[k].fd = fd;
        job[k].offset = len * k / n;
        job[k].bytes = len * (k + 1) / n - job[k].offset;
        job[k].dst = dst + job[k].offset;
        job[k].got = 0;
    }
    for (k = 1; k < n; ++k) {
        if (pthread_create(&thread[k], NULL, slice_read, &job[k]) != 0) break;
        ++started;
    }
    slice_read(&job[0]);
    for (k = started + 1; k < n; ++k) slice_read(&job[k]);
    for (k = 1; k <= started; ++k) pthread_join(thread[k], NULL);
    for (k = 0; k < n; ++k) {
        if (job[k].got != job[k].bytes) return -1; /* short read (not a regular file?): let fread sort it out */
        total += job[k].got;
    }
    return total;
}

/* the whole stream in one read (the reference measures the file with fseek/ftell too, src/util.c:20-34) */
static struct shared_text *slurp(FILE *f)
{
    struct shared_text *t = calloc(1, sizeof *t);
    long got;
    if (!t) {
        MEM_FAIL();
        return NULL;
    }
    t->refs = 1;
    if (fseek(f, 0, SEEK_END) != 0 || (t->len = ftell(f)) < 0) {
        /* not seekable (a pipe: `cat x.pdb | program`).  The reference's ftell()-based range logic degenerates to "read
         * lines until EOF" there (src/util.c:20-34, src/structure.c:658); do the same: everything up to EOF */
        size_t cap = 1 << 20, len = 0, n;
        clearerr(f);
        if (!(t->data = malloc(cap + 1))) {
            MEM_FAIL();
            goto fail;
        }
        while ((n = fread(t->data + len, 1, cap - len, f)) > 0) {
            len += n;
            if (len == cap) {
                char *p = realloc(t->data, (cap *= 2) + 1);
                if (!p) {
                    MEM_FAIL();
                    goto fail;
                }
                t->data = p;
            }
        }
        if (ferror(f)) {
            FAIL_MSG("%s", strerror(errno));
            goto fail;
        }
        t->len = (long)len;
        t->data[len] = '\0';
        return t;
    }
    rewind(f);
    if (!(t->data = malloc((size_t)t->len + 1))) {
        MEM_FAIL();
        goto fail;
    }
    got = -1;
    if (t->len >= SLURP_PARALLEL_MIN_BYTES) got = slurp_parallel(f, t->data, t->len);
    if (got < 0) { /* small file, or the descriptor could not be read in slices: one fread */
        rewind(f);
        got = (long)fread(t->data, 1, (size_t)t->len, f);
        if (ferror(f)) {
            FAIL_MSG("%s", strerror(errno));
            goto fail;
        }
    } else {
        fseek(f, 0, SEEK_END); /* where a full fread leaves the stream */
    }
    t->len = got;
    t->data[got] = '\0';
    return t;
fail:
    fsb_text_release(t);
    return NULL;
}

/* One fgets(line, 120, f) of the reference: up to 119 bytes, through the first newline.  *raw = bytes consumed,
 * return value = strlen() of what the reference would see (an embedded NUL cuts the string). */
static inline long next_line(const struct shared_text *t, long pos, long *raw)
{
    const long room = t->len - pos < LINE_MAX_STRL - 1 ? t->len - pos : LINE_MAX_STRL - 1;
    const char *p = t->data + pos, *nl = memchr(p, '\n', (size_t)room), *nul;
    const long n = nl ? nl - p + 1 : room;
    *raw = n;
    nul = memchr(p, '\0', (size_t)n);
    return nul ? nul - p : n;
}

static inline int is_atom_line(const char *l, long len, int options)
{
    if (len >= 4 && memcmp(l, "ATOM", 4) == 0) return 1;
    return (options & FREESASA_INCLUDE_HETATM) && len >= 6 && memcmp(l, "HETATM", 6) == 0;
}

/* freesasa_pdb_ishydrogen(), src/pdb.c:260-283 (the caller has established that the line is an ATOM/HETATM record,
 * which is what pdb_line_check() re-verifies).  Non-zero = treated as hydrogen; lines shorter than 13 characters
 * give FREESASA_FAIL there, which the caller also reads as "true". */
static inline int is_hydrogen(const char *l, long len)
{
    if (len < 13) return 1;
    if (len >= 78) { /* element symbol present, columns 77-78 */
        if (l[76] == ' ' && (l[77] == 'H' || l[77] == 'D')) return 1;
        if (!(l[76] == ' ' && l[77] == ' ')) return 0;
    } else {
        return 0; /* symbol "" is "not blank and not H or D" */
    }
    if (!(l[12] == ' ' || (l[12] >= '1' && l[12] <= '9'))) return 0;
    if (l[12] == 'H' || l[13] == 'H') return 1;
    if (l[12] == 'D' || l[13] == 'D') return 1;
    return 0;
}

static inline int is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

/* The layout nearly every PDB file uses: three right-justified %8.3f fields.  Accepted only when reading the section
 * as whitespace-separated tokens (what the reference does) must give the same three numbers: each field is blanks, an
 * optional '-', 1-4 digits, '.', three digits, and fields two and three do not start with a digit (otherwise the previous
 * token would run on).  value = integer / 1000.0, one correctly rounded division. */
static inline int strict_coords(const char *s, double v[3])
{
    int k;
    for (k = 0; k < 3; ++k, s += 8) {
        unsigned d5 = (unsigned)(s[5] - '0'), d6 = (unsigned)(s[6] - '0'), d7 = (unsigned)(s[7] - '0'), d3 = (unsigned)(s[3] - '0');
        int m, j, neg = 0;
        if (s[4] != '.' || d3 > 9 || d5 > 9 || d6 > 9 || d7 > 9) return 0;
        m = (int)d3;
        for (j = 2; j >= 0; --j) {
            const unsigned d = (unsigned)(s[j] - '0');
            if (d > 9) break;
            m += (int)d * (j == 2 ? 10 : j == 1 ? 100 : 1000);
        }
        if (j >= 0) {
            if (s[j] == '-') {
                neg = 1;
                --j;
            }
            for (; j >= 0; --j)
                if (s[j] != ' ') return 0;
        } else if (k > 0) {
            return 0; /* a digit in the first column of a later field: the tokens would merge */
        }
        m = m * 1000 + (int)(d5 * 100 + d6 * 10 + d7);
        v[k] = neg ? -((double)m / 1000.0) : (double)m / 1000.0;
    }
    return 1;
}

/* Three whitespace-separated decimal numbers from the 24-character coordinate section.  Fast path: sign, at most 15
 * digits with an optional point -> integer mantissa / 10^k, one IEEE division, which is the correctly rounded value
 * strtod()/sscanf("%lf") return.  Anything else (exponent, hex, inf, nan, long mantissa, missing number) -> 0 and the
 * caller repeats the reference's own sscanf. */
static int fast_coords(const char *s, int n, double v[3])
{
    static const double pow10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
    int i = 0, k;
    for (k = 0; k < 3; ++k) {
        uint64_t m = 0;
        int neg = 0, digits = 0, frac = 0, seen_point = 0;
        while (i < n && is_space(s[i])) ++i;
        if (i < n && (s[i] == '-' || s[i] == '+')) neg = s[i++] == '-';
        for (; i < n; ++i) {
            const char c = s[i];
            if (c >= '0' && c <= '9') {
                m = m * 10 + (uint64_t)(c - '0');
                ++digits;
                frac += seen_point;
            } else if (c == '.' && !seen_point) {
                seen_point = 1;
            } else {
                break;
            }
        }
        if (digits == 0 || digits > 15) return 0;
        if (i < n && (s[i] == 'e' || s[i] == 'E' || s[i] == 'x' || s[i] == 'X' || s[i] == 'p' || s[i] == 'P')) return 0;
        v[k] = frac ? (double)m / pow10[frac] : (double)m;
        if (neg) v[k] = -v[k];
    }
    return 1;
}

/* Direct-mapped memo of classifier lookups keyed by the raw 4+3 label bytes of a line: a structure has a few hundred
 * distinct (residue, atom) pairs, so after the first residues every lookup is one load and one compare. */
#define MEMO_SLOTS 1024
struct memo {
    uint64_t key[MEMO_SLOTS];
    int row[MEMO_SLOTS];
};
static inline int memo_row(struct memo *m, const freesasa_classifier *classifier, const struct atom_label *a)
{
    uint64_t key = 0;
    uint32_t slot;
    memcpy(&key, a->name, 4);
    memcpy((char *)&key + 4, a->res_name, 3);
    slot = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 54);
    if (m->key[slot] != key || key == 0) {
        m->key[slot] = key;
        m->row[slot] = fsb_classifier_row(classifier, a->res_name, a->name);
    }
    return m->row[slot];
}

/* What a pass over a byte range saw besides atoms. */
struct scan_info {
    int saw_model, model; /* last MODEL record (only looked at without FREESASA_JOIN_MODELS) */
    int hit_endmdl;       /* stopped at an ENDMDL record */
    int odd_lines;        /* a line longer than the reference's fgets() buffer, or one with an embedded NUL */
};

/* The line loop of from_pdb_impl(), src/structure.c:658-707, over [begin, end] of the text, appending to `s`.  Starts
 * with "no alternate location seen".  Returns FREESASA_SUCCESS or FREESASA_FAIL (message already reported). */
static int parse_lines(freesasa_structure *s, const struct shared_text *t, long begin, long end,
                       const freesasa_classifier *classifier, int options, struct memo *memo, struct scan_info *info)
{
    const freesasa_classifier *cl = classifier ? classifier : &freesasa_default_classifier;
    char the_alt = ' ';
    long pos = begin, raw, len;

    while (pos < t->len) {
        const char *l = t->data + pos;
        len = next_line(t, pos, &raw);
        pos += raw;
        if (pos > end) break;
        if (len != raw || (raw == LINE_MAX_STRL - 1 && l[raw - 1] != '\n')) info->odd_lines = 1;

        if (is_atom_line(l, len, options)) {
            struct atom_label a;
            char alt;
            double v[3];
            int ret;

            if (is_hydrogen(l, len) && !(options & FREESASA_INCLUDE_HYDROGEN)) continue;

            /* atom_new_from_line(), src/structure.c:201-240: a field is empty when the line is too short for it */
            alt = len > 16 ? l[16] : '\0'; /* line[16] of a 16-character line is its terminator */
            memset(&a, 0, sizeof a);
            if (len >= 16) memcpy(a.name, l + 12, 4);
            if (len >= 20) memcpy(a.res_name, l + 17, 3);
            if (len >= 27) memcpy(a.res_number, l + 22, 5);
            a.chain[0] = len > 21 ? l[21] : '\0';
            if (len >= 78) memcpy(a.symbol, l + 76, 2);
            if (len < 78 || (a.symbol[0] == ' ' && a.symbol[1] == ' ')) guess_symbol(a.symbol, a.name);

            /* only the first alternate location of a run is kept (src/structure.c:672-679) */
            if ((alt != ' ' && the_alt == ' ') || alt == ' ')
                the_alt = alt;
            else if (alt != ' ' && alt != the_alt)
                continue;

            /* freesasa_pdb_get_coord(), src/pdb.c:176-197 */
            if (len < 54) return FREESASA_FAIL;
            if (!strict_coords(l + 30, v) && !fast_coords(l + 30, 24, v)) {
                char section[25];
                memcpy(section, l + 30, 24);
                section[24] = '\0';
                if (sscanf(section, "%lf%lf%lf", &v[0], &v[1], &v[2]) != 3) {
                    char copy[LINE_MAX_STRL];
                    memcpy(copy, l, (size_t)len);
                    copy[len] = '\0';
                    return FAIL_MSG("could not read coordinates from line '%s'", copy);
                }
            }

            ret = add_atom(s, &a, v, classifier, options, l, (size_t)len, memo_row(memo, cl, &a));
            if (ret == FREESASA_FAIL) return FREESASA_FAIL;
            if (ret == FREESASA_WARN) continue;

            if (options & FREESASA_RADIUS_FROM_OCCUPANCY) {
                /* freesasa_pdb_get_occupancy() -> pdb_get_double(line + 54, 6), src/pdb.c:33-49,239-247 */
                char buf[8];
                float occ;
                long w = len - 54 < 6 ? len - 54 : 6;
                if (len < 55) return FREESASA_FAIL;
                memcpy(buf, l + 54, (size_t)w);
                buf[w] = '\0';
                if (sscanf(buf, "%f", &occ) != 1) return FREESASA_FAIL;
                s->radius[s->n - 1] = occ;
            }
        }

        if (!(options & FREESASA_JOIN_MODELS)) {
            if (len >= 5 && memcmp(l, "MODEL", 5) == 0 && len > 10) {
                char rest[LINE_MAX_STRL];
                memcpy(rest, l + 10, (size_t)(len - 10));
                rest[len - 10] = '\0';
                if (sscanf(rest, "%d", &info->model) == 1) info->saw_model = 1;
            }
            if (len >= 6 && memcmp(l, "ENDMDL", 6) == 0) {
                info->hit_endmdl = 1;
                break;
            }
        }
    }
    return FREESASA_SUCCESS;
}

static freesasa_structure *fragment_new(const struct shared_text *t, long span)
{
    freesasa_structure *s = freesasa_structure_new();
    if (s == NULL) return NULL;
    s->text = (struct shared_text *)t;
    __atomic_add_fetch(&s->text->refs, 1, __ATOMIC_RELAXED);
    /* an atom needs a line of at least 54 characters: size every array once, no reallocation while reading (untouched
     * capacity costs address space only) */
    if (span > 0 && reserve_atoms(s, (int)(span / 54) + 1)) {
        freesasa_structure_free(s);
        return NULL;
    }
    return s;
}

/* ---- parallel reading -------------------------------------------------------------------------------------
 * A large range is cut at line boundaries into one chunk per thread; every chunk is parsed into a private fragment
 * (same code as the serial reader, messages captured), then the fragments are concatenated in order.  What makes the
 * result IDENTICAL to a serial pass:
 *   - "first alternate location wins" carries state from line to line: a chunk may only start right after an atom
 *     line that is kept and has a blank alternate-location column, where that state is known to be "none";
 *   - residues and chains are re-derived at the seams with the serial rules (a residue continues across a seam when
 *     number and chain agree; a chain label is registered the first time it appears);
 *   - MODEL numbers: the last MODEL record wins; ENDMDL: every chunk after the first one that met it is discarded;
 *   - messages are replayed in chunk order up to the first failure;
 *   - anything odd (over-long lines, NUL bytes, no safe cut found) falls back to the serial reader.
 */
#define PARALLEL_MIN_BYTES (1L << 20)
#define PARALLEL_MAX_THREADS 16

struct chunk_job {
    const struct shared_text *t;
    long begin, end;
    const freesasa_classifier *classifier;
    int options;
    freesasa_structure *frag;
    struct scan_info info;
    struct fsb_capture messages;
    int status;
    /* phase 2 */
    freesasa_structure *dst;
    int atom_offset, res_offset;
};

static long n_parallel_reads; /* test hook: how many ranges went through the parallel reader */
long fsb_ingest_parallel_reads(void) { return __atomic_load_n(&n_parallel_reads, __ATOMIC_RELAXED); }

static void *chunk_parse(void *arg)
{
    struct chunk_job *j = arg;
    struct memo *memo = calloc(1, sizeof *memo);
    fsb_capture_current = &j->messages;
    j->frag = fragment_new(j->t, j->end - j->begin);
    if (j->frag == NULL || memo == NULL)
        j->status = FREESASA_FAIL;
    else
        j->status = parse_lines(j->frag, j->t, j->begin, j->end, j->classifier, j->options, memo, &j->info);
    fsb_capture_current = NULL;
    free(memo);
    return NULL;
}

static void *chunk_copy(void *arg)
{
    struct chunk_job *j = arg;
    const freesasa_structure *f = j->frag;
    freesasa_structure *d = j->dst;
    const int o = j->atom_offset, n = f->n;
    int i;
    memcpy(d->coord.xyz + 3 * (size_t)o, f->coord.xyz, 3 * (size_t)n * sizeof(double));
    memcpy(d->radius + o, f->radius, (size_t)n * sizeof(double));
    memcpy(d->label + o, f->label, (size_t)n * sizeof(struct atom_label));
    memcpy(d->cls + o, f->cls, (size_t)n);
    memcpy(d->line_at + o, f->line_at, (size_t)n * sizeof(long));
    memcpy(d->line_len + o, f->line_len, (size_t)n);
    for (i = 0; i < n; ++i) d->res_index[o + i] = f->res_index[i] + j->res_offset;
    return NULL;
}


/* First position >= from where a chunk may start (see above), or -1 if none before `limit`. */
static long safe_cut(const struct shared_text *t, long from, long limit, int options)
{
    long pos = from, raw, len;
    const char *nl = memchr(t->data + pos, '\n', (size_t)(limit - pos));
    int budget = 4096;
    if (!nl) return -1;
    pos = nl - t->data + 1; /* a line start */
    while (pos < limit && budget-- > 0) {
        const char *l = t->data + pos;
        len = next_line(t, pos, &raw);
        pos += raw;
        if (len != raw || (raw == LINE_MAX_STRL - 1 && l[raw - 1] != '\n')) return -1;
        if (is_atom_line(l, len, options) && !(is_hydrogen(l, len) && !(options & FREESASA_INCLUDE_HYDROGEN)) && len > 16 &&
            l[16] == ' ')
            return pos;
    }
    return -1;
}

/* Returns 1 and sets *out (NULL on a reported failure) if the range was read in parallel, 0 to ask for the serial reader. */
static int from_range_parallel(const struct shared_text *t, long begin, long end, const freesasa_classifier *classifier,
                               int options, freesasa_structure **out)
{
    struct chunk_job job[PARALLEL_MAX_THREADS];
    pthread_t thread[PARALLEL_MAX_THREADS];
    const long stop = end < t->len ? end : t->len, span = stop - begin;
    int n_threads = fsb_hardware_threads(), n_jobs = 0, k, used, failed = -1, started;
    freesasa_structure *s = NULL;
    long cut = begin;

    {
        /* FREESASA_B200_PARALLEL_MIN_BYTES: test hook, lets the test-suite drive small inputs through this path */
        const char *env = getenv("FREESASA_B200_PARALLEL_MIN_BYTES");
        const long min_bytes = env ? atol(env) : PARALLEL_MIN_BYTES, per_thread = min_bytes / 4 > 256 ? min_bytes / 4 : 256;
        if (span < min_bytes || n_threads < 2) return 0;
        if (n_threads > span / per_thread) n_threads = (int)(span / per_thread);
    }
    memset(job, 0, sizeof job);
    for (k = 0; k < n_threads; ++k) {
        long next = k == n_threads - 1 ? stop : safe_cut(t, begin + span * (k + 1) / n_threads, stop, options);
        if (next < 0) next = stop; /* no safe cut: this chunk takes the rest */
        if (next <= cut) continue;
        job[n_jobs].t = t;
        job[n_jobs].begin = cut;
        /* the serial loop stops at the first line that ENDS beyond `end`; chunk ends are line starts, so the last byte
         * of a chunk's last line is next - 1 */
        job[n_jobs].end = next == stop ? end : next;
        job[n_jobs].classifier = classifier;
        job[n_jobs].options = options;
        ++n_jobs;
        cut = next;
        if (cut >= stop) break;
    }
    if (n_jobs < 2) return 0;

    started = 0;
    for (k = 1; k < n_jobs; ++k) {
        if (pthread_create(&thread[k], NULL, chunk_parse, &job[k]) != 0) break;
        ++started;
    }
    chunk_parse(&job[0]);
    for (k = started + 1; k < n_jobs; ++k) chunk_parse(&job[k]); /* threads that could not be created: do it here */
    for (k = 1; k <= started; ++k) pthread_join(thread[k], NULL);

    __atomic_add_fetch(&n_parallel_reads, 1, __ATOMIC_RELAXED);
    /* how many chunks count: up to the first failure or ENDMDL; odd lines anywhere -> serial reader decides */
    used = n_jobs;
    for (k = 0; k < n_jobs; ++k) {
        if (job[k].info.odd_lines) goto serial;
        if (job[k].status != FREESASA_SUCCESS) {
            failed = k;
            used = k + 1;
            break;
        }
        if (job[k].info.hit_endmdl) {
            used = k + 1;
            break;
        }
    }
    for (k = 0; k < used; ++k) fsb_capture_flush(&job[k].messages);
    *out = NULL;
    if (failed < 0) {
        int total = 0, n_res = 0;
        for (k = 0; k < used; ++k) total += job[k].frag->n;
        if (total == 0) {
            FAIL_MSG("input had no valid ATOM or HETATM lines");
            failed = used;
        } else if ((s = fragment_new(t, 0)) == NULL || reserve_atoms(s, total)) {
            failed = used;
        } else {
            /* seams: residue continuation, chain registration, offsets */
            const struct atom_label *prev = NULL;
            int atoms = 0;
            for (k = 0; k < used; ++k) {
                const freesasa_structure *f = job[k].frag;
                int r, c, merged;
                if (f->classifier_name && !s->classifier_name) {
                    s->classifier_name = strdup(f->classifier_name);
                    s->last_classifier = f->last_classifier;
                }
                if (job[k].info.saw_model) s->model = job[k].info.model;
                if (f->n == 0) continue;
                merged = prev && !(memcmp(f->label[0].res_number, prev->res_number, sizeof prev->res_number) ||
                                   memcmp(f->label[0].chain, prev->chain, sizeof prev->chain));
                job[k].dst = s;
                job[k].atom_offset = atoms;
                job[k].res_offset = n_res - (merged ? 1 : 0);
                if (n_res + f->n_res > s->cap_res) {
                    const int cap = 2 * (n_res + f->n_res) + 256;
                    void *p;
                    if (!(p = realloc(s->res_first, (size_t)cap * sizeof(int)))) goto nomem;
                    s->res_first = p;
                    if (!(p = realloc(s->res_ref, (size_t)cap * sizeof(freesasa_nodearea)))) goto nomem;
                    s->res_ref = p;
                    if (!(p = realloc(s->res_has_ref, (size_t)cap))) goto nomem;
                    s->res_has_ref = p;
                    s->cap_res = cap;
                }
                for (r = merged ? 1 : 0; r < f->n_res; ++r) {
                    s->res_first[n_res] = f->res_first[r] + atoms;
                    s->res_ref[n_res] = f->res_ref[r];
                    s->res_has_ref[n_res] = f->res_has_ref[r];
                    ++n_res;
                }
                for (c = 0; c < f->n_chains; ++c) {
                    int known = 0, q;
                    for (q = 0; q < s->n_chains && !known; ++q) known = memcmp(s->chain_label[q], f->chain_label[c], 4) == 0;
                    if (known) continue;
                    if (s->n_chains == s->cap_chains) {
                        const int cap = s->cap_chains ? 2 * s->cap_chains : 64;
                        void *p;
                        if (!(p = realloc(s->chain_label, (size_t)cap * 4))) goto nomem;
                        s->chain_label = p;
                        if (!(p = realloc(s->short_labels, (size_t)cap + 1))) goto nomem;
                        s->short_labels = p;
                        if (!(p = realloc(s->chain_first, (size_t)cap * sizeof(int)))) goto nomem;
                        s->chain_first = p;
                        s->cap_chains = cap;
                    }
                    memcpy(s->chain_label[s->n_chains], f->chain_label[c], 4);
                    s->short_labels[s->n_chains] = f->chain_label[c][0];
                    s->short_labels[s->n_chains + 1] = '\0';
                    s->chain_first[s->n_chains] = f->chain_first[c] + atoms;
                    ++s->n_chains;
                }
                prev = &f->label[f->n - 1];
                atoms += f->n;
            }
            s->n_res = n_res;
            s->n = s->coord.n = total;
            /* phase 2: every fragment is copied into place by its own thread (first touch of the final arrays in parallel) */
            {
                int spawned[PARALLEL_MAX_THREADS] = {0};
                for (k = 1; k < used; ++k) {
                    if (job[k].dst == NULL) continue;
                    spawned[k] = pthread_create(&thread[k], NULL, chunk_copy, &job[k]) == 0;
                    if (!spawned[k]) chunk_copy(&job[k]);
                }
                if (job[0].dst) chunk_copy(&job[0]);
                for (k = 1; k < used; ++k)
                    if (spawned[k]) pthread_join(thread[k], NULL);
            }
        }
    }
    if (failed >= 0) {
        FAIL_MSG("%s", "");
        freesasa_structure_free(s);
        s = NULL;
    }
    for (k = 0; k < n_jobs; ++k) {
        free(job[k].messages.text);
        freesasa_structure_free(job[k].frag);
    }
    *out = s;
    return 1;
nomem:
    MEM_FAIL();
    failed = used;
    FAIL_MSG("%s", "");
    freesasa_structure_free(s);
    for (k = 0; k < n_jobs; ++k) {
        free(job[k].messages.text);
        freesasa_structure_free(job[k].frag);
    }
    *out = NULL;
    return 1;
serial:
    for (k = 0; k < n_jobs; ++k) {
        free(job[k].messages.text);
        freesasa_structure_free(job[k].frag);
    }
    return 0;
}

/* from_pdb_impl(), src/structure.c:638-721, on the byte range [begin, end] of the text */
static freesasa_structure *from_range_opt(const struct shared_text *t, long begin, long end, const freesasa_classifier *classifier,
                                          int options, int allow_threads)
{
    freesasa_structure *s = NULL;
    struct memo *memo;
    struct scan_info info = {0, 0, 0, 0};

    if (allow_threads && from_range_parallel(t, begin, end, classifier, options, &s)) return s;
    s = fragment_new(t, (end < t->len ? end : t->len) - begin);
    memo = calloc(1, sizeof *memo);
    if (s == NULL || memo == NULL) {
        free(memo);
        freesasa_structure_free(s);
        return NULL;
    }
    if (parse_lines(s, t, begin, end, classifier, options, memo, &info)) goto fail;
    if (info.saw_model) s->model = info.model;
    if (s->n == 0) {
        FAIL_MSG("input had no valid ATOM or HETATM lines");
        goto fail;
    }
    free(memo);
    return s;
fail:
    FAIL_MSG("%s", "");
    free(memo);
    freesasa_structure_free(s);
    return NULL;
}

static freesasa_structure *from_range(const struct shared_text *t, long begin, long end, const freesasa_classifier *classifier,
                                      int options)
{
    return from_range_opt(t, begin, end, classifier, options, 1);
}

freesasa_structure *freesasa_structure_from_pdb(FILE *pdb_file, const freesasa_classifier *classifier, int options)
{
    struct shared_text *t;
    freesasa_structure *s;
    assert(pdb_file);
    if (!(t = slurp(pdb_file))) return NULL;
    s = from_range(t, 0, t->len, classifier, options);
    fsb_text_release(t);
    return s;
}

/* Additive: the same reader on a memory buffer (what a caller that already holds the file bytes would use).  The
 * bytes are copied once; the caller's buffer is not referenced after the call. */
freesasa_structure *freesasa_structure_from_pdb_buffer(const char *text, long len, const freesasa_classifier *classifier,
                                                       int options)
{
    struct shared_text *t = calloc(1, sizeof *t);
    freesasa_structure *s;
    assert(text);
    if (!t || !(t->data = malloc((size_t)len + 1))) {
        free(t);
        MEM_FAIL();
        return NULL;
    }
    memcpy(t->data, text, (size_t)len);
    t->data[len] = '\0';
    t->len = len;
    t->refs = 1;
    s = from_range(t, 0, len, classifier, options);
    fsb_text_release(t);
    return s;
}

/* ---- models and chains as separate structures (freesasa_structure_array(), src/structure.c:848-953) ------------ */
struct range {
    long begin, end;
};

/* freesasa_pdb_get_models(), src/pdb.c:51-98 */
static int find_models(const struct shared_text *t, struct range **out)
{
    struct range *m = NULL;
    int n = 0, n_end = 0, cap = 0;
    long pos = 0, raw, len;
    while (pos < t->len) {
        const char *l = t->data + pos;
        len = next_line(t, pos, &raw);
        if (len >= 5 && memcmp(l, "MODEL", 5) == 0) {
            if (n == cap) {
                void *p = realloc(m, sizeof(struct range) * (size_t)(cap = 2 * cap + 8));
                if (!p) {
                    free(m);
                    *out = NULL;
                    return MEM_FAIL();
                }
                m = p;
            }
            m[n].begin = pos;
            m[n].end = t->len; /* a MODEL without its ENDMDL (truncated file) runs to the end of the text; the reference
                                * leaves this field uninitialised (src/pdb.c:63-73) and does whatever the heap held */
            ++n;
        }
        if (len >= 6 && memcmp(l, "ENDMDL", 6) == 0) {
            ++n_end;
            if (n != n_end) {
                free(m);
                *out = NULL;
                return FAIL_MSG("mismatch between MODEL and ENDMDL in input");
            }
            m[n - 1].end = pos + raw;
        }
        pos += raw;
    }
    if (n == 0) {
        free(m);
        m = NULL;
    }
    *out = m;
    return n;
}

/* freesasa_pdb_get_chains(), src/pdb.c:100-150: a new range starts whenever the chain column of an atom line differs
 * from the previous atom line's; only lines that end strictly before model.end are looked at */
static int find_chains(const struct shared_text *t, struct range model, struct range **out, int options)
{
    struct range *c = NULL;
    int n = 0, cap = 0;
    char last_chain = '\0';
    long pos = model.begin, last_pos = model.begin, raw, len;
    *out = NULL;
    while (pos < t->len) {
        const char *l = t->data + pos;
        len = next_line(t, pos, &raw);
        if (!(pos + raw < model.end)) break;
        if (is_atom_line(l, len, options)) {
            const char chain = len > 21 ? l[21] : '\0';
            if (chain != last_chain) {
                if (n > 0) c[n - 1].end = last_pos;
                if (n == cap) {
                    void *p = realloc(c, sizeof(struct range) * (size_t)(cap = 2 * cap + 8));
                    if (!p) {
                        free(c);
                        return MEM_FAIL();
                    }
                    c = p;
                }
                c[n++].begin = last_pos;
                last_chain = chain;
            }
        }
        pos += raw;
        last_pos = pos;
    }
    if (n > 0) {
        c[n - 1].end = last_pos;
        c[0].begin = model.begin; /* keeps the MODEL record */
        *out = c;
    }
    return n;
}

/* One structure-to-be of freesasa_structure_array(): a byte range and the model number it gets, or (range.begin < 0) just
 * a message that the serial reader would have printed at this point. */
struct piece {
    struct range range;
    int model;
    freesasa_structure *structure;
    struct fsb_capture messages;
};
struct piece_work {
    const struct shared_text *t;
    const freesasa_classifier *classifier;
    int options, n_pieces, next;
    struct piece *pieces;
};
static void piece_worker(int part, int n_parts, void *arg)
{
    struct piece_work *w = arg;
    (void)part;
    (void)n_parts;
    for (;;) {
        const int k = __atomic_fetch_add(&w->next, 1, __ATOMIC_RELAXED);
        struct piece *p;
        if (k >= w->n_pieces) break;
        p = &w->pieces[k];
        if (p->range.begin < 0) continue;
        fsb_capture_current = &p->messages;
        p->structure = from_range_opt(w->t, p->range.begin, p->range.end, w->classifier, w->options, 0);
        fsb_capture_current = NULL;
    }
}

/* freesasa_structure_array(), src/structure.c:848-953.  The ranges (models, chains) are found first; with many of them
 * (an NMR ensemble, the chains of a large assembly) they are parsed concurrently, each with its messages captured, and
 * the outcome — structures, messages, the point of failure — is replayed in file order, i.e. exactly the serial one. */
freesasa_structure **freesasa_structure_array(FILE *pdb, int *n, const freesasa_classifier *classifier, int options)
{
    struct shared_text *t;
    struct range *models = NULL, *chains = NULL, whole;
    struct piece *pieces = NULL;
    freesasa_structure **ss = NULL;
    int n_models, n_pieces = 0, cap_pieces = 0, n_total = 0, i, j, failed = 0, threads;

    assert(pdb);
    assert(n);
    *n = 0;
    if (!((options & FREESASA_SEPARATE_MODELS) || (options & FREESASA_SEPARATE_CHAINS))) {
        FAIL_MSG("options need to specify at least one of FREESASA_SEPARATE_CHAINS and FREESASA_SEPARATE_MODELS");
        return NULL;
    }
    if (!(t = slurp(pdb))) return NULL;
    whole.begin = 0;
    whole.end = t->len;
    n_models = find_models(t, &models);
    if (n_models == FREESASA_FAIL) {
        FAIL_MSG("problems reading PDB-file");
        fsb_text_release(t);
        return NULL;
    }
    if (n_models == 0) {
        models = &whole;
        n_models = 1;
    }
    if (!(options & FREESASA_SEPARATE_MODELS)) n_models = 1; /* only the first model */

    /* the pieces, in file order */
    for (i = 0; i < n_models && !failed; ++i) {
        int n_new = 1;
        if (options & FREESASA_SEPARATE_CHAINS) {
            n_new = find_chains(t, models[i], &chains, options);
            if (n_new == FREESASA_FAIL) {
                failed = 1;
                break;
            }
        }
        if (n_pieces + (n_new > 0 ? n_new : 1) > cap_pieces) {
            void *p = realloc(pieces, sizeof(struct piece) * (size_t)(cap_pieces = 2 * cap_pieces + n_new + 16));
            if (!p) {
                MEM_FAIL();
                failed = 1;
                break;
            }
            pieces = p;
        }
        if (n_new == 0) { /* a model without atoms: the reference warns and goes on */
            struct piece *p = &pieces[n_pieces++];
            memset(p, 0, sizeof *p);
            p->range.begin = -1;
            fsb_capture_current = &p->messages;
            WARN_MSG("in %s(): no chains found (in model %d)", __func__, i + 1);
            fsb_capture_current = NULL;
            continue;
        }
        for (j = 0; j < n_new; ++j) {
            struct piece *p = &pieces[n_pieces++];
            memset(p, 0, sizeof *p);
            p->range = (options & FREESASA_SEPARATE_CHAINS) ? chains[j] : models[i];
            p->model = i + 1;
        }
        free(chains);
        chains = NULL;
    }

    /* parse: concurrently when there are enough pieces to go round, else one after the other (where a large piece may
     * itself be read in parallel); stop at the first failure either way */
    threads = fsb_hardware_threads();
    if (!failed && n_pieces >= 4 && threads >= 2) {
        struct piece_work w = {t, classifier, options, n_pieces, 0, pieces};
        fsb_parallel_run(threads < n_pieces ? threads : n_pieces, piece_worker, &w);
    } else if (!failed) {
        for (i = 0; i < n_pieces; ++i) {
            fsb_capture_flush(&pieces[i].messages); /* a "no chains found" warning, at its place in the sequence */
            if (pieces[i].range.begin < 0) continue;
            pieces[i].structure = from_range(t, pieces[i].range.begin, pieces[i].range.end, classifier, options);
            if (pieces[i].structure == NULL) break; /* later pieces are never read */
        }
    }
    /* replay in file order */
    if (!failed && n_pieces > 0 && !(ss = calloc((size_t)n_pieces, sizeof(freesasa_structure *)))) {
        MEM_FAIL();
        failed = 1;
    }
    for (i = 0; i < n_pieces && !failed; ++i) {
        fsb_capture_flush(&pieces[i].messages);
        if (pieces[i].range.begin < 0) continue;
        if (pieces[i].structure == NULL) {
            failed = 1;
            break;
        }
        pieces[i].structure->model = pieces[i].model;
        ss[n_total++] = pieces[i].structure;
        pieces[i].structure = NULL;
    }
    if (n_total == 0) failed = 1;
    for (i = 0; i < n_pieces; ++i) { /* whatever was not handed over */
        free(pieces[i].messages.text);
        freesasa_structure_free(pieces[i].structure);
    }
    free(pieces);
    free(chains);
    if (models != &whole) free(models);
    fsb_text_release(t);
    if (failed) {
        for (i = 0; i < n_total; ++i) freesasa_structure_free(ss[i]);
        free(ss);
        return NULL;
    }
    *n = n_total;
    return ss;
}

/* freesasa_structure_get_chains() / _lcl(), src/structure.c:955-1081 */
static int in_group(const freesasa_chain_group *group, const char *chain)
{
    size_t k;
    for (k = 0; k < group->n; ++k)
        if (strncmp(group->chains[k], chain, 4) == 0) return 1;
    return 0;
}
static freesasa_structure *get_chains(const freesasa_structure *structure, const char *chains, const freesasa_chain_group *group,
                                      const freesasa_classifier *classifier, int options)
{
    freesasa_structure *out;
    int i;
    if (!(out = freesasa_structure_new())) return NULL;
    out->model = structure->model;
    for (i = 0; i < structure->n; ++i) {
        const struct atom_label *a = &structure->label[i];
        if (group ? in_group(group, a->chain) : strchr(chains, a->chain[0]) != NULL) {
            const double *v = structure->coord.xyz + 3 * i;
            if (add_atom_wopt(out, a->name, a->res_name, a->res_number, a->symbol, a->chain, v[0], v[1], v[2], classifier,
                              options) == FREESASA_FAIL) {
                FAIL_MSG("%s", "");
                goto fail;
            }
        }
    }
    if (out->n == 0) goto fail;
    if ((size_t)out->n_chains != (group ? group->n : strlen(chains))) {
        FAIL_MSG("structure has chains '%s', but '%s' requested", structure->short_labels, group ? "(chain group)" : chains);
        goto fail;
    }
    return out;
fail:
    freesasa_structure_free(out);
    return NULL;
}
freesasa_structure *freesasa_structure_get_chains(const freesasa_structure *structure, const char *chains,
                                                  const freesasa_classifier *classifier, int options)
{
    assert(structure);
    if (strlen(chains) == 0) return NULL;
    return get_chains(structure, chains, NULL, classifier, options);
}
freesasa_structure *freesasa_structure_get_chains_lcl(const freesasa_structure *structure, const freesasa_chain_group *chains,
                                                      const freesasa_classifier *classifier, int options)
{
    assert(structure);
    assert(chains);
    if (chains->n == 0) return NULL;
    return get_chains(structure, NULL, chains, classifier, options);
}

/* ---- accessors (src/structure.c:1083-1424) ------------------------------------------------------------ */
#define ATOM_OK(s, i) (assert(s), assert((i) >= 0 && (i) < (s)->n))
#define RES_OK(s, r) (assert(s), assert((r) >= 0 && (r) < (s)->n_res))

const char *freesasa_structure_chain_labels(const freesasa_structure *s) { return s->short_labels; }
int freesasa_structure_n(const freesasa_structure *s) { return s->n; }
int freesasa_structure_n_residues(const freesasa_structure *s) { return s->n_res; }
int freesasa_structure_n_chains(const freesasa_structure *s) { return s->n_chains; }
const double *freesasa_structure_radius(const freesasa_structure *s) { return s->radius; }
void freesasa_structure_set_radius(freesasa_structure *s, const double *radii)
{
    assert(s);
    assert(radii);
    memcpy(s->radius, radii, (size_t)s->n * sizeof(double));
}
const char *freesasa_structure_atom_name(const freesasa_structure *s, int i) { ATOM_OK(s, i); return s->label[i].name; }
const char *freesasa_structure_atom_res_name(const freesasa_structure *s, int i) { ATOM_OK(s, i); return s->label[i].res_name; }
const char *freesasa_structure_atom_res_number(const freesasa_structure *s, int i) { ATOM_OK(s, i); return s->label[i].res_number; }
char freesasa_structure_atom_chain(const freesasa_structure *s, int i) { ATOM_OK(s, i); return s->label[i].chain[0]; }
const char *freesasa_structure_atom_chain_lcl(const freesasa_structure *s, int i) { ATOM_OK(s, i); return s->label[i].chain; }
const char *freesasa_structure_atom_symbol(const freesasa_structure *s, int i) { ATOM_OK(s, i); return s->label[i].symbol; }
double freesasa_structure_atom_radius(const freesasa_structure *s, int i) { ATOM_OK(s, i); return s->radius[i]; }
void freesasa_structure_atom_set_radius(freesasa_structure *s, int i, double radius) { ATOM_OK(s, i); s->radius[i] = radius; }
freesasa_atom_class freesasa_structure_atom_class(const freesasa_structure *s, int i) { ATOM_OK(s, i); return (freesasa_atom_class)FSB_CLS_CLASS(s->cls[i]); }
/* src/structure.c:1199-1206: the line as fgets() delivered it (with its newline), NULL for atoms added by hand.
 * The NUL-terminated copies are built for the whole structure on the first call. */
const char *freesasa_structure_atom_pdb_line(const freesasa_structure *s, int i)
{
    ATOM_OK(s, i);
    if (s->line_at[i] < 0) return NULL;
    if (__atomic_load_n(&s->lines, __ATOMIC_ACQUIRE) == NULL) {
        static pthread_mutex_t once = PTHREAD_MUTEX_INITIALIZER;
        pthread_mutex_lock(&once);
        if (s->lines == NULL) {
            char *all = malloc((size_t)s->n * LINE_MAX_STRL);
            int k;
            if (all == NULL) {
                pthread_mutex_unlock(&once);
                MEM_FAIL();
                return NULL;
            }
            for (k = 0; k < s->n; ++k) {
                if (s->line_at[k] >= 0) memcpy(all + (size_t)k * LINE_MAX_STRL, s->text->data + s->line_at[k], s->line_len[k]);
                all[(size_t)k * LINE_MAX_STRL + (s->line_at[k] >= 0 ? s->line_len[k] : 0)] = '\0';
            }
            __atomic_store_n((char **)&s->lines, all, __ATOMIC_RELEASE);
        }
        pthread_mutex_unlock(&once);
    }
    return s->lines + (size_t)i * LINE_MAX_STRL;
}
int fsb_structure_atom_residue(const freesasa_structure *s, int i) { ATOM_OK(s, i); return s->res_index[i]; }

const char *freesasa_structure_residue_name(const freesasa_structure *s, int r) { RES_OK(s, r); return s->label[s->res_first[r]].res_name; }
const char *freesasa_structure_residue_number(const freesasa_structure *s, int r) { RES_OK(s, r); return s->label[s->res_first[r]].res_number; }
char freesasa_structure_residue_chain(const freesasa_structure *s, int r) { RES_OK(s, r); return s->label[s->res_first[r]].chain[0]; }
const freesasa_nodearea *freesasa_structure_residue_reference(const freesasa_structure *s, int r)
{
    RES_OK(s, r);
    return s->res_has_ref[r] ? &s->res_ref[r] : NULL;
}
int freesasa_structure_residue_atoms(const freesasa_structure *s, int r, int *first, int *last)
{
    RES_OK(s, r);
    assert(first);
    assert(last);
    *first = s->res_first[r];
    *last = r == s->n_res - 1 ? s->n - 1 : s->res_first[r + 1] - 1;
    return FREESASA_SUCCESS;
}

static int chain_index(const freesasa_structure *s, const char *chain)
{
    int i;
    for (i = 0; i < s->n_chains; ++i)
        if (strncmp(s->chain_label[i], chain, 4) == 0) return i;
    return FAIL_MSG("chain '%s' not found", chain);
}
/* src/structure.c:1303-1325: a chain's atoms run up to the first atom of the next registered chain */
int freesasa_structure_chain_atoms_lcl(const freesasa_structure *s, const char *chain, int *first, int *last)
{
    int c;
    assert(s);
    if ((c = chain_index(s, chain)) < 0) return FAIL_MSG("%s", "");
    *first = s->chain_first[c];
    *last = c == s->n_chains - 1 ? s->n - 1 : s->chain_first[c + 1] - 1;
    return FREESASA_SUCCESS;
}
int freesasa_structure_chain_atoms(const freesasa_structure *s, char chain, int *first, int *last)
{
    const char label[2] = {chain, '\0'};
    return freesasa_structure_chain_atoms_lcl(s, label, first, last);
}
int freesasa_structure_chain_residues_lcl(const freesasa_structure *s, const char *chain, int *first, int *last)
{
    int fa, la;
    assert(s);
    if (freesasa_structure_chain_atoms_lcl(s, chain, &fa, &la)) return FAIL_MSG("%s", "");
    *first = s->res_index[fa];
    *last = s->res_index[la];
    return FREESASA_SUCCESS;
}
int freesasa_structure_chain_residues(const freesasa_structure *s, char chain, int *first, int *last)
{
    const char label[2] = {chain, '\0'};
    return freesasa_structure_chain_residues_lcl(s, label, first, last);
}
const char *freesasa_structure_residue_chain_lcl(const freesasa_structure *s, int r)
{
    RES_OK(s, r);
    return s->label[s->res_first[r]].chain;
}
const char *freesasa_structure_chain_label(const freesasa_structure *s, int index)
{
    assert(s);
    assert(index >= 0 && index < s->n_chains);
    return s->chain_label[index];
}
int freesasa_structure_model(const freesasa_structure *s) { return s->model; }
void freesasa_structure_set_model(freesasa_structure *s, int model) { s->model = model; }
const char *freesasa_structure_classifier_name(const freesasa_structure *s) { return s->classifier_name; }
const double *freesasa_structure_coord_array(const freesasa_structure *s) { return s->coord.xyz; }
const coord_t *freesasa_structure_xyz(const freesasa_structure *s) { return &s->coord; }

/* ---- into the engine ------------------------------------------------------------------------------------ */
/* src/freesasa.c:144-153 */
freesasa_result *freesasa_calc_structure(const freesasa_structure *structure, const freesasa_parameters *parameters)
{
    assert(structure);
    return freesasa_calc(freesasa_structure_xyz(structure), freesasa_structure_radius(structure), parameters);
}

/* Row f-2: what the CLI's loop over structures (src/main.cc:334-362: one freesasa_calc_tree per model / chain group)
 * becomes with a batched engine: every structure of the array in one device pass. */
int freesasa_calc_structure_batch(int n_struct, freesasa_structure *const *structures, const freesasa_parameters *parameters,
                                  freesasa_result **results)
{
    const double **xyz, **radii;
    int *n_atoms, k, rc;
    if (n_struct <= 0 || !structures || !results) return FAIL_MSG("invalid batch arguments");
    xyz = malloc(sizeof(double *) * (size_t)n_struct);
    radii = malloc(sizeof(double *) * (size_t)n_struct);
    n_atoms = malloc(sizeof(int) * (size_t)n_struct);
    if (!xyz || !radii || !n_atoms) {
        rc = MEM_FAIL();
    } else {
        for (k = 0; k < n_struct; ++k) {
            assert(structures[k]);
            xyz[k] = structures[k]->coord.xyz;
            radii[k] = structures[k]->radius;
            n_atoms[k] = structures[k]->n;
        }
        rc = freesasa_calc_coord_batch(n_struct, xyz, radii, n_atoms, parameters, results);
    }
    free(xyz);
    free(radii);
    free(n_atoms);
    return rc;
}
