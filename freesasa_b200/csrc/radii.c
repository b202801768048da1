/* freesasa_b200/csrc/radii.c — atom classifiers: (residue, atom) -> radius / polarity class.
 *
 * Scope row f-1 of SURVEY.md §8 ("PDB ingest + classifier lookup").  Mirrors the public classifier
 * API of the reference (src/freesasa.h:403-607, src/classifier.c) with the same lookup rules:
 *
 *   - keys are compared after trimming to their first whitespace-delimited token
 *     (find_string(), src/classifier.c:126-160);
 *   - an atom is looked up in its own residue first, then in the pseudo-residue "ANY"
 *     (find_atom()/find_any(), src/classifier.c:739-779);
 *   - unknown atoms give radius -1.0 / class FREESASA_ATOM_UNKNOWN (src/classifier.c:781-811).
 *
 * What is different is the machinery: the reference does a malloc + sscanf + linear strcmp scan over
 * residues and then atoms, three times per atom read (radius, class, residue reference;
 * src/structure.c:498-541,605-625).  Here every classifier is a flat row table with two open-addressing
 * hash indexes (pair -> row, residue -> residue row) built once; a lookup is one FNV hash of <= 8 bytes
 * and normally one probe, no allocation.
 *
 * The built-in tables (ProtOr, NACCESS, OONS) and the element radii are data extracted from the compiled
 * reference by tests/golden/make_radius_tables.py into radius_tables.inc.
 */
#include "host_internal.h"

#include <assert.h>
#include <ctype.h>
#include <errno.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

struct radius_row {
    const char *residue, *atom;
    double radius;
    int cls;
};
struct residue_row {
    const char *name;
    int has_ref; /* 0: the reference stores a nodearea with a NULL name and zeros */
    double total, main_chain, side_chain, polar, apolar, unknown;
};
struct element_row {
    char symbol[3];
    double radius;
};
#include "radius_tables.inc"

/* hash indexes of one classifier, built once by index_build() (the built-in classifiers are const objects, as in
 * the reference's header, so their index lives in a separate mutable block) */
struct lookup {
    uint32_t *pair_slot; /* row + 1 */
    uint32_t pair_mask;
    uint32_t *res_slot; /* residue + 1 */
    uint32_t res_mask;
    freesasa_nodearea *ref; /* one per residue, what freesasa_classifier_residue_reference() hands out */
};
struct freesasa_classifier {
    const char *name;
    const struct radius_row *rows;
    int n_rows;
    const struct residue_row *residues;
    int n_residues;
    struct lookup *index;
    int owned; /* heap-allocated by freesasa_classifier_from_file() */
};

#define COUNT(a) ((int)(sizeof(a) / sizeof((a)[0])))
static struct lookup protor_lookup, naccess_lookup, oons_lookup;
const freesasa_classifier freesasa_protor_classifier = {protor_name, protor_rows, COUNT(protor_rows), protor_residues, COUNT(protor_residues), &protor_lookup, 0};
const freesasa_classifier freesasa_naccess_classifier = {naccess_name, naccess_rows, COUNT(naccess_rows), naccess_residues, COUNT(naccess_residues), &naccess_lookup, 0};
const freesasa_classifier freesasa_oons_classifier = {oons_name, oons_rows, COUNT(oons_rows), oons_residues, COUNT(oons_residues), &oons_lookup, 0};

/* ---- hashing ------------------------------------------------------------------------------------ */
static inline uint32_t fnv(uint32_t h, const char *s, int len)
{
    int i;
    for (i = 0; i < len; ++i) h = (h ^ (unsigned char)s[i]) * 16777619u;
    return h;
}
static inline uint32_t pair_hash(const char *res, int rl, const char *atom, int al)
{
    uint32_t h = fnv(fnv(2166136261u, res, rl) * 16777619u ^ 0x2fu, atom, al);
    return h ^ (h >> 15);
}
static inline uint32_t res_hash(const char *res, int rl)
{
    uint32_t h = fnv(2166136261u, res, rl);
    return h ^ (h >> 15);
}
static inline int same(const char *stored, const char *key, int len) { return strncmp(stored, key, (size_t)len) == 0 && stored[len] == '\0'; }

static uint32_t pow2_at_least(int n)
{
    uint32_t p = 16;
    while (p < (uint32_t)n) p <<= 1;
    return p;
}

static int index_build(const freesasa_classifier *cl)
{
    int i;
    uint32_t h;
    struct lookup *c = cl->index;
    const uint32_t np = pow2_at_least(2 * cl->n_rows + 1), nr = pow2_at_least(2 * cl->n_residues + 1);
    c->pair_slot = calloc(np, sizeof(uint32_t));
    c->res_slot = calloc(nr, sizeof(uint32_t));
    c->ref = calloc((size_t)(cl->n_residues > 0 ? cl->n_residues : 1), sizeof(freesasa_nodearea));
    if (!c->pair_slot || !c->res_slot || !c->ref) return MEM_FAIL();
    c->pair_mask = np - 1;
    c->res_mask = nr - 1;
    for (i = 0; i < cl->n_residues; ++i) {
        const struct residue_row *r = &cl->residues[i];
        const int len = (int)strlen(r->name);
        freesasa_nodearea a = {r->has_ref ? r->name : NULL, r->total, r->main_chain, r->side_chain, r->polar, r->apolar, r->unknown};
        c->ref[i] = a;
        for (h = res_hash(r->name, len) & c->res_mask; c->res_slot[h]; h = (h + 1) & c->res_mask)
            if (same(cl->residues[c->res_slot[h] - 1].name, r->name, len)) break; /* first one wins */
        if (!c->res_slot[h]) c->res_slot[h] = (uint32_t)i + 1;
    }
    for (i = 0; i < cl->n_rows; ++i) {
        const struct radius_row *r = &cl->rows[i];
        const int rl = (int)strlen(r->residue), al = (int)strlen(r->atom);
        for (h = pair_hash(r->residue, rl, r->atom, al) & c->pair_mask; c->pair_slot[h]; h = (h + 1) & c->pair_mask) {
            const struct radius_row *o = &cl->rows[c->pair_slot[h] - 1];
            if (same(o->residue, r->residue, rl) && same(o->atom, r->atom, al)) break;
        }
        if (!c->pair_slot[h]) c->pair_slot[h] = (uint32_t)i + 1;
    }
    return FREESASA_SUCCESS;
}

__attribute__((constructor)) static void builtin_indexes(void)
{
    if (index_build(&freesasa_protor_classifier) || index_build(&freesasa_naccess_classifier) ||
        index_build(&freesasa_oons_classifier))
        abort();
}

static int pair_find(const freesasa_classifier *cl, const char *res, int rl, const char *atom, int al)
{
    const struct lookup *c = cl->index;
    uint32_t h;
    if (rl == 0 || al == 0) return -1; /* the reference compares uninitialised memory here; nothing can match */
    for (h = pair_hash(res, rl, atom, al) & c->pair_mask; c->pair_slot[h]; h = (h + 1) & c->pair_mask) {
        const struct radius_row *o = &cl->rows[c->pair_slot[h] - 1];
        if (same(o->residue, res, rl) && same(o->atom, atom, al)) return (int)c->pair_slot[h] - 1;
    }
    return -1;
}

/* Row of (res_name, atom_name) with the ANY fallback of src/classifier.c:756-779, or -1. */
int fsb_classifier_row(const freesasa_classifier *c, const char *res_name, const char *atom_name)
{
    const char *r, *a;
    const int rl = fsb_token(res_name, &r), al = fsb_token(atom_name, &a);
    int row = pair_find(c, r, rl, a, al);
    if (row < 0) row = pair_find(c, "ANY", 3, a, al);
    return row;
}
double fsb_classifier_row_radius(const freesasa_classifier *c, int row) { return c->rows[row].radius; }
int fsb_classifier_row_class(const freesasa_classifier *c, int row) { return c->rows[row].cls; }

static int residue_find(const freesasa_classifier *cl, const char *res_name)
{
    const struct lookup *c = cl->index;
    const char *r;
    const int rl = fsb_token(res_name, &r);
    uint32_t h;
    if (rl == 0) return -1;
    for (h = res_hash(r, rl) & c->res_mask; c->res_slot[h]; h = (h + 1) & c->res_mask)
        if (same(cl->residues[c->res_slot[h] - 1].name, r, rl)) return (int)c->res_slot[h] - 1;
    return -1;
}

/* ---- public API (reference src/freesasa.h:545-607) --------------------------------------------------- */
double freesasa_classifier_radius(const freesasa_classifier *classifier, const char *res_name, const char *atom_name)
{
    int row;
    assert(classifier);
    assert(res_name);
    assert(atom_name);
    row = fsb_classifier_row(classifier, res_name, atom_name);
    return row >= 0 ? classifier->rows[row].radius : -1.0;
}

freesasa_atom_class freesasa_classifier_class(const freesasa_classifier *classifier, const char *res_name,
                                              const char *atom_name)
{
    int row;
    assert(classifier);
    assert(res_name);
    assert(atom_name);
    row = fsb_classifier_row(classifier, res_name, atom_name);
    return row >= 0 ? (freesasa_atom_class)classifier->rows[row].cls : FREESASA_ATOM_UNKNOWN;
}

const char *freesasa_classifier_class2str(freesasa_atom_class atom_class)
{
    switch (atom_class) {
    case FREESASA_ATOM_APOLAR: return "Apolar";
    case FREESASA_ATOM_POLAR: return "Polar";
    case FREESASA_ATOM_UNKNOWN: return "Unknown";
    }
    FAIL_MSG("invalid atom class");
    return NULL;
}

const char *freesasa_classifier_name(const freesasa_classifier *classifier) { return classifier->name; }

/* src/classifier.c:853-862: non-NULL whenever the residue is known to the classifier (the area's name is
 * NULL and its values 0 when the classifier has no reference values for it) */
const freesasa_nodearea *freesasa_classifier_residue_reference(const freesasa_classifier *classifier, const char *res_name)
{
    const int res = residue_find(classifier, res_name);
    return res < 0 ? NULL : &classifier->index->ref[res];
}

/* src/classifier.c:1002-1017: the symbol is right-justified to two characters and compared exactly */
double freesasa_guess_radius(const char *input_symbol)
{
    char symbol[3];
    int i;
    assert(input_symbol);
    snprintf(symbol, sizeof symbol, "%2s", input_symbol);
    for (i = 0; i < COUNT(element_rows); ++i)
        if (symbol[0] == element_rows[i].symbol[0] && symbol[1] == element_rows[i].symbol[1]) return element_rows[i].radius;
    return -1.0;
}

/* src/classifier.c:1090-1108 */
int freesasa_atom_is_backbone(const char *atom_name)
{
    static const char *bb[] = {"CA", "N", "O", "C", "OXT", "P", "OP1", "OP2", "O5'", "C5'", "C4'", "O4'", "C3'", "O3'", "C2'", "C1'"};
    const char *a;
    const int al = fsb_token(atom_name, &a);
    int i;
    if (al == 0) return 0;
    for (i = 0; i < COUNT(bb); ++i)
        if (same(bb[i], a, al)) return 1;
    return 0;
}

/* Residue types for the per-type sums of freesasa_write_res() (src/classifier.c:1020-1088): the twenty amino acids first
 * (always listed), then non-standard ones, "UNK" (everything not recognised), capping groups and nucleotides. */
static const char *const residue_type_name[] = {
    "ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE", "LEU", "LYS", "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL",
    "CSE", "SEC", "PYL", "PYH", "ASX", "GLX", "UNK", "ACE", "NH2", "DA", "DC", "DG", "DT", "DU", "DI", "A", "C", "G", "U", "I", "T", "N"};
#define RESIDUE_TYPE_UNKNOWN 26
int freesasa_classify_n_residue_types(void) { return COUNT(residue_type_name); }
const char *freesasa_classify_residue_name(int residue_type)
{
    assert(residue_type >= 0 && residue_type < COUNT(residue_type_name));
    return residue_type_name[residue_type];
}
int freesasa_classify_residue(const char *res_name)
{
    const char *r;
    const int len = fsb_token(res_name, &r);
    int i;
    for (i = 0; i < COUNT(residue_type_name); ++i)
        if (len > 0 && same(residue_type_name[i], r, len)) return i;
    return RESIDUE_TYPE_UNKNOWN;
}

/* ---- user configuration files (src/classifier.c:164-735; format: doc/doxy-main.md "Classifier configuration") -- */
#define CFG_LINE 256 /* MAX_LINE_LEN, src/classifier.c:18 */

struct cfg_type {
    char *name;
    double radius;
    int cls;
};
struct cfg {
    char *text;
    long len;
    struct cfg_type *types;
    int n_types;
    struct radius_row *rows;
    int n_rows, cap_rows;
    struct residue_row *residues;
    int n_residues, cap_residues;
};

/* one fgets(buf, CFG_LINE + 1) worth of text starting at pos; returns its length (0 at the end) */
static long cfg_chunk(const struct cfg *c, long pos)
{
    long n = 0;
    while (pos + n < c->len && n < CFG_LINE) {
        ++n;
        if (c->text[pos + n - 1] == '\n') break;
    }
    return n;
}

/* strip_line(), src/classifier.c:163-195: cut at '#', trim blanks/tabs on the left and blanks/tabs/newlines on
 * the right; a result of fewer than two characters counts as empty (the reference's `first >= last`). */
static int cfg_strip(char *out, const char *in, long n)
{
    long first = 0, last;
    const char *hash = memchr(in, '#', (size_t)n);
    const char *nul = memchr(in, '\0', (size_t)n);
    if (nul) n = nul - in;
    if (hash && hash - in < n) n = hash - in;
    last = n - 1;
    while (first < n && (in[first] == ' ' || in[first] == '\t')) ++first;
    if (last > first)
        while (in[last] == ' ' || in[last] == '\t' || in[last] == '\n') --last;
    if (first >= last) {
        out[0] = '\0';
        return 0;
    }
    memcpy(out, in + first, (size_t)(last - first + 1));
    out[last - first + 1] = '\0';
    return (int)(last - first + 1);
}

static int parse_class(const char *name)
{
    if (strncasecmp(name, "apolar", 6) == 0) return FREESASA_ATOM_APOLAR;
    if (strncasecmp(name, "polar", 5) == 0) return FREESASA_ATOM_POLAR;
    return FAIL_MSG("only atom classes allowed are 'polar' and 'apolar' (case insensitive)");
}

static int cfg_find_type(const struct cfg *c, const char *name)
{
    int i;
    for (i = 0; i < c->n_types; ++i)
        if (strcmp(c->types[i].name, name) == 0) return i;
    return -1;
}

static int cfg_types_line(struct cfg *c, const char *line)
{
    char t[CFG_LINE + 1], cl[CFG_LINE + 1];
    double r;
    int cls;
    void *p;
    if (sscanf(line, "%s %lf %s", t, &r, cl) != 3)
        return FAIL_MSG("could not parse line '%s' in configuration, expecting triplet of type 'TYPE [RADIUS] CLASS' for "
                        "example 'C_ALI 2.00 apolar'", line);
    if (cfg_find_type(c, t) >= 0) return WARN_MSG("ignoring duplicate configuration entry for '%s'", t);
    cls = parse_class(cl);
    if (cls == FREESASA_FAIL) return FAIL_MSG("%s", "");
    p = realloc(c->types, sizeof(struct cfg_type) * (size_t)(c->n_types + 1));
    if (!p) return MEM_FAIL();
    c->types = p;
    if (!(c->types[c->n_types].name = strdup(t))) return MEM_FAIL();
    c->types[c->n_types].radius = r;
    c->types[c->n_types].cls = cls;
    ++c->n_types;
    return FREESASA_SUCCESS;
}

static int cfg_atoms_line(struct cfg *c, const char *line)
{
    char res[CFG_LINE + 1], atom[CFG_LINE + 1], type[CFG_LINE + 1];
    int t, i, known = 0;
    if (sscanf(line, "%s %s %s", res, atom, type) != 3)
        return FAIL_MSG("could not parse configuration, line '%s', expecting triplet of type 'RESIDUE ATOM CLASS', for "
                        "example 'ALA CB C_ALI'", line);
    if (strlen(res) > 3) return FAIL_MSG("residue name %s is too long in classifier file", res);
    if (strlen(atom) > 4) return FAIL_MSG("atom name %s is too long in classifier file", atom);
    if ((t = cfg_find_type(c, type)) < 0) return FAIL_MSG("unknown atom type '%s' in configuration, line '%s'", type, line);
    for (i = 0; i < c->n_residues && !known; ++i) known = strcmp(c->residues[i].name, res) == 0;
    if (!known) {
        if (c->n_residues == c->cap_residues) {
            void *p = realloc(c->residues, sizeof(struct residue_row) * (size_t)(c->cap_residues = 2 * c->cap_residues + 16));
            if (!p) return MEM_FAIL();
            c->residues = p;
        }
        memset(&c->residues[c->n_residues], 0, sizeof(struct residue_row));
        if (!(c->residues[c->n_residues].name = strdup(res))) return MEM_FAIL();
        ++c->n_residues;
    }
    for (i = 0; i < c->n_rows; ++i)
        if (strcmp(c->rows[i].residue, res) == 0 && strcmp(c->rows[i].atom, atom) == 0)
            return WARN_MSG("ignoring duplicate configuration entry for atom '%s %s'", res, atom);
    if (c->n_rows == c->cap_rows) {
        void *p = realloc(c->rows, sizeof(struct radius_row) * (size_t)(c->cap_rows = 2 * c->cap_rows + 64));
        if (!p) return MEM_FAIL();
        c->rows = p;
    }
    c->rows[c->n_rows].residue = strdup(res);
    c->rows[c->n_rows].atom = strdup(atom);
    if (!c->rows[c->n_rows].residue || !c->rows[c->n_rows].atom) return MEM_FAIL();
    c->rows[c->n_rows].radius = c->types[t].radius;
    c->rows[c->n_rows].cls = c->types[t].cls;
    ++c->n_rows;
    return FREESASA_SUCCESS;
}

/* read_types()/read_atoms(), src/classifier.c:476-515,644-673: the first line of the range is the section keyword
 * and is discarded; the status of the LAST non-empty line is the status of the section (so a trailing duplicate,
 * which is only a warning, still makes the reference reject the file — kept). */
static int cfg_section(struct cfg *c, long begin, long end, int (*one)(struct cfg *, const char *))
{
    char line[CFG_LINE + 1];
    long pos = begin, n;
    int ret = FREESASA_SUCCESS;
    n = cfg_chunk(c, pos);
    if (n == 0 || cfg_strip(line, c->text + pos, n) <= 0) return FREESASA_FAIL;
    pos += n;
    while (pos < end) {
        n = cfg_chunk(c, pos);
        if (n == 0) break;
        if (cfg_strip(line, c->text + pos, n) > 0) {
            ret = one(c, line);
            if (ret == FREESASA_FAIL) break;
        }
        pos += n;
    }
    return ret;
}

static void cfg_release(struct cfg *c, int keep_tables)
{
    int i;
    for (i = 0; i < c->n_types; ++i) free(c->types[i].name);
    free(c->types);
    free(c->text);
    if (!keep_tables) {
        for (i = 0; i < c->n_rows; ++i) {
            free((char *)c->rows[i].residue);
            free((char *)c->rows[i].atom);
        }
        for (i = 0; i < c->n_residues; ++i) free((char *)c->residues[i].name);
        free(c->rows);
        free(c->residues);
    }
}

freesasa_classifier *freesasa_classifier_from_file(FILE *file)
{
    struct cfg c;
    struct range {
        long begin, end;
    } types = {-1, -1}, atoms = {-1, -1}, name = {-1, -1}, *prev = NULL, *which[3];
    static const char *keyword[3] = {"types:", "atoms:", "name:"};
    long pos, n, start, got;
    char line[CFG_LINE + 1], *cname = NULL;
    freesasa_classifier *out = NULL;
    int k;

    assert(file);
    memset(&c, 0, sizeof c);
    which[0] = &types, which[1] = &atoms, which[2] = &name;

    /* slurp from the current position (the reference scans from ftell() to EOF, src/classifier.c:316-347) */
    start = ftell(file);
    if (start < 0 || fseek(file, 0, SEEK_END) != 0 || (c.len = ftell(file) - start) < 0 || fseek(file, start, SEEK_SET) != 0) {
        FAIL_MSG("%s", strerror(errno));
        goto fail;
    }
    if (!(c.text = malloc((size_t)c.len + 1))) {
        MEM_FAIL();
        goto fail;
    }
    got = (long)fread(c.text, 1, (size_t)c.len, file);
    if (ferror(file)) {
        FAIL_MSG("%s", strerror(errno));
        goto fail;
    }
    c.len = got;
    c.text[c.len] = '\0';
    rewind(file);

    /* check_file(): locate the sections */
    for (pos = 0; (n = cfg_chunk(&c, pos)) > 0; pos += n) {
        long visible = n;
        const char *hash = memchr(c.text + pos, '#', (size_t)n), *nul = memchr(c.text + pos, '\0', (size_t)n);
        if (nul) visible = nul - (c.text + pos);
        if (hash && hash - (c.text + pos) < visible) visible = hash - (c.text + pos);
        memcpy(line, c.text + pos, (size_t)visible);
        line[visible] = '\0';
        for (k = 0; k < 3 && visible > 0; ++k) {
            const char *loc = strstr(line, keyword[k]);
            if (loc) {
                which[k]->begin = pos + (loc - line);
                if (prev) prev->end = which[k]->begin;
                prev = which[k];
            }
        }
        if (n == CFG_LINE && c.text[pos + n - 1] != '\n') {
            FAIL_MSG("Lines in classifier files can only be %d characters or less", CFG_LINE);
            goto fail;
        }
    }
    if (prev) prev->end = c.len;
    if (name.begin == -1) WARN_MSG("input configuration lacks the entry 'name:', will use 'no-name-given'");
    if (types.begin == -1 || atoms.begin == -1) {
        FAIL_MSG("input configuration lacks (at least) one of the entries 'types:' or 'atoms:'");
        goto fail;
    }

    /* read_name(), src/classifier.c:675-700: the token after "name:" on the same line */
    if (name.begin >= 0) {
        char tok[CFG_LINE + 1] = "";
        pos = name.begin + 5;
        n = cfg_chunk(&c, pos);
        memcpy(line, c.text + pos, (size_t)n);
        line[n] = '\0';
        sscanf(line, "%s", tok);
        if (tok[0] == '\0') {
            FAIL_MSG("empty name for configuration?");
            goto fail;
        }
        cname = strdup(tok);
    } else {
        cname = strdup("no-name-given");
    }
    if (!cname) {
        MEM_FAIL();
        goto fail;
    }

    if (cfg_section(&c, types.begin, types.end, cfg_types_line)) goto fail;
    if (cfg_section(&c, atoms.begin, atoms.end, cfg_atoms_line)) goto fail;

    if (!(out = calloc(1, sizeof *out)) || !(out->index = calloc(1, sizeof *out->index))) {
        free(out);
        MEM_FAIL();
        goto fail;
    }
    out->name = cname;
    out->rows = c.rows;
    out->n_rows = c.n_rows;
    out->residues = c.residues;
    out->n_residues = c.n_residues;
    out->owned = 1;
    cfg_release(&c, 1);
    if (index_build(out)) {
        freesasa_classifier_free(out);
        FAIL_MSG("%s", "");
        return NULL;
    }
    return out;
fail:
    free(cname);
    cfg_release(&c, 0);
    FAIL_MSG("%s", "");
    return NULL;
}

void freesasa_classifier_free(freesasa_classifier *c)
{
    int i;
    if (c == NULL || !c->owned) return; /* the built-in classifiers are static */
    for (i = 0; i < c->n_rows; ++i) {
        free((char *)c->rows[i].residue);
        free((char *)c->rows[i].atom);
    }
    for (i = 0; i < c->n_residues; ++i) free((char *)c->residues[i].name);
    free((void *)c->rows);
    free((void *)c->residues);
    free((char *)c->name);
    free(c->index->pair_slot);
    free(c->index->res_slot);
    free(c->index->ref);
    free(c->index);
    free(c);
}
