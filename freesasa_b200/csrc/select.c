/* freesasa_b200/csrc/select.c — selections: "name, expression" -> the area of the selected atoms.
 *
 * Scope row f-4 of SURVEY.md §8(f).  Mirrors freesasa_selection_new() / freesasa_select_area() and friends
 * (reference src/freesasa.h:610-690,1855-1882, src/selection.c) with the reference's language:
 *
 *   stmt   :  NAME ',' expr
 *   expr   :  '(' expr ')' | expr and expr | expr or expr | not expr
 *          |  resn list | symbol list | name list | resi r_range | chain c_range
 *   list   :  id ('+' id)*            r_range : items joined by '+', item = id | id '-' id | '-' id | id '-'
 *   id     :  NUMBER | ID | '\-' NUMBER           c_range : items joined by '+', item = id | id '-' id
 *
 * The reference builds an expression tree with flex + bison and evaluates it with one int per atom and node; here a
 * hand-written scanner (the longest-match rules of src/lexer.l, trailing context for the selection name included) feeds a
 * precedence-climbing parser (the %left/%precedence table of src/parser.y: or < and < not < '+' < '-') that evaluates
 * directly into byte masks.  The semantics that matter for parity are kept to the letter: identifiers are upper-cased,
 * labels are compared by their first whitespace-delimited token, "resi" ranges use atoi() of the residue number (so
 * insertion codes fall inside the range of their number), open ranges run from the first / to the last atom's number,
 * invalid identifiers are warned about and ignored, and the area is the serial sum over ALL atoms of flag x area, which
 * is bit-identical to the reference's because adding 0.0 changes nothing.
 */
#include "host_internal.h"

#include <assert.h>
#include <ctype.h>
#include <stdlib.h>
#include <string.h>

#define MAX_SELECTION_NAME 50 /* FREESASA_MAX_SELECTION_NAME, src/freesasa.h:226 */

struct freesasa_selection {
    char *name;
    char *command;
    double area;
    int n_atoms;
};

/* ---- scanner (src/lexer.l) ---------------------------------------------------------------------------------- */
enum token { T_END, T_ERROR, T_COMMA, T_DASH, T_PLUS, T_LPAR, T_RPAR, T_RESN, T_RESI, T_SYMBOL, T_NAME, T_CHAIN, T_AND, T_OR, T_NOT,
             T_MINUS, T_NUMBER, T_ID, T_SELID };

struct scanner {
    const char *p;
    enum token tok;
    char text[256]; /* NUMBER, ID, SELID */
};

static int is_word(int c) { return isalnum(c) || c == '_'; }

static int keyword(const char *s, int len, const char *kw)
{
    int i;
    if ((int)strlen(kw) != len) return 0;
    for (i = 0; i < len; ++i)
        if (tolower((unsigned char)s[i]) != kw[i]) return 0;
    return 1;
}

/* One token, flex style: the longest match wins, the earlier rule on a tie.  The candidates at a position are a
 * punctuation mark, a keyword/number/identifier over [alnum_]+'* and the selection name [alnum_+-]+ followed by ','
 * (whose match length counts the comma, which is how "s1," beats the identifier "s1"). */
static void next_token(struct scanner *sc)
{
    const char *p = sc->p;
    int word = 0, selid = 0, all_digits = 1, len;
    while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') ++p;
    sc->text[0] = '\0';
    if (*p == '\0') {
        sc->p = p;
        sc->tok = T_END;
        return;
    }
    while (is_word((unsigned char)p[word])) {
        if (!isdigit((unsigned char)p[word])) all_digits = 0;
        ++word;
    }
    if (word > 0 && !all_digits) /* ID: [alnum_]+ followed by any number of apostrophes (C5', O3'); NUMBER has none */
        while (p[word] == '\'') ++word;
    else if (word > 0 && p[word] == '\'') { /* digits then an apostrophe: the ID rule matches more than NUMBER */
        all_digits = 0;
        while (p[word] == '\'') ++word;
    }
    while (is_word((unsigned char)p[selid]) || p[selid] == '-' || p[selid] == '+') ++selid;
    if (selid > 0 && p[selid] == ',' && selid + 1 > word && selid + 1 > 1) {
        len = selid < (int)sizeof sc->text - 1 ? selid : (int)sizeof sc->text - 1;
        memcpy(sc->text, p, (size_t)len);
        sc->text[len] = '\0';
        sc->p = p + selid;
        sc->tok = T_SELID;
        return;
    }
    if (word > 0) {
        /* a keyword only if the whole word is the keyword (otherwise ID is the longer match) */
        sc->p = p + word;
        if (keyword(p, word, "resn")) sc->tok = T_RESN;
        else if (keyword(p, word, "resi")) sc->tok = T_RESI;
        else if (keyword(p, word, "symbol")) sc->tok = T_SYMBOL;
        else if (keyword(p, word, "name")) sc->tok = T_NAME;
        else if (keyword(p, word, "chain")) sc->tok = T_CHAIN;
        else if (keyword(p, word, "and")) sc->tok = T_AND;
        else if (keyword(p, word, "or")) sc->tok = T_OR;
        else if (keyword(p, word, "not")) sc->tok = T_NOT;
        else {
            len = word < (int)sizeof sc->text - 1 ? word : (int)sizeof sc->text - 1;
            memcpy(sc->text, p, (size_t)len);
            sc->text[len] = '\0';
            sc->tok = all_digits ? T_NUMBER : T_ID;
        }
        return;
    }
    sc->p = p + 1;
    switch (*p) {
    case ',': sc->tok = T_COMMA; return;
    case '-': sc->tok = T_DASH; return;
    case '+': sc->tok = T_PLUS; return;
    case '(': sc->tok = T_LPAR; return;
    case ')': sc->tok = T_RPAR; return;
    case '&': sc->tok = T_AND; return;
    case '|': sc->tok = T_OR; return;
    case '!': sc->tok = T_NOT; return;
    case '\\':
        if (p[1] == '-') {
            sc->p = p + 2;
            sc->tok = T_MINUS;
            return;
        }
        break;
    }
    sc->tok = T_ERROR; /* a character no rule knows */
}

/* ---- evaluation (src/selection.c:267-560) --------------------------------------------------------------------- */
enum selector { S_NAME, S_SYMBOL, S_RESN, S_RESI, S_CHAIN };
static const char *selector_str[] = {"name", "symbol", "resn", "resi", "chain"};

struct ident {
    int is_number; /* E_NUMBER (digits, possibly with a leading '-' from "\-5") or E_ID */
    char value[256];
};

struct parse {
    struct scanner sc;
    const freesasa_structure *s;
    int n;
    int dry; /* syntax check only: the reference parses the whole command before it evaluates (and warns about) anything */
    int depth; /* nesting of parentheses and "not": bounded, so that a hostile command cannot exhaust the stack */
    int warn, err;
};

static int token_equals(const char *field, const char *id)
{
    const char *t;
    const int len = fsb_token(field, &t);
    return (int)strlen(id) == len && memcmp(t, id, (size_t)len) == 0;
}

/* select_id(): flag every atom whose field matches; warn when nothing does */
static void select_id(struct parse *ps, enum selector sel, const char *id, unsigned char *mask)
{
    const freesasa_structure *s = ps->s;
    int i, count = 0;
    for (i = 0; i < ps->n; ++i) {
        int match = 0;
        switch (sel) {
        case S_NAME: match = token_equals(s->label[i].name, id); break;
        case S_SYMBOL: match = token_equals(s->label[i].symbol, id); break;
        case S_RESN: match = token_equals(s->label[i].res_name, id); break;
        case S_RESI: match = token_equals(s->label[i].res_number, id); break;
        case S_CHAIN: match = id[0] == s->label[i].chain[0]; break;
        }
        if (match) mask[i] = 1;
        count += match;
    }
    if (count == 0) WARN_MSG("Found no matches to %s '%s', typo?", selector_str[sel], id);
}

/* is_valid_id(), src/selection.c:375-447 */
static int valid_id(enum selector sel, const struct ident *id)
{
    const char *val = id->value;
    const int n = (int)strlen(val);
    int warn = 0, i;
    switch (sel) {
    case S_NAME:
        if (n > 4) return WARN_MSG("select: %s: atom name '%s' invalid (string too long), will be ignored", selector_str[sel], val);
        break;
    case S_SYMBOL:
        if (id->is_number)
            return WARN_MSG("select: %s: '%s' invalid (should be 1 or 2 letters, 'C', 'N', 'SE', etc), will be ignored", selector_str[sel], val);
        if (n > 2) return WARN_MSG("select: %s: '%s' invalid (element names have 1 or 2 characters), will be ignored", selector_str[sel], val);
        break;
    case S_RESN:
        if (n > 3) return WARN_MSG("select: %s: '%s' invalid (string too long), will be ignored", selector_str[sel], val);
        break;
    case S_RESI:
        if (!id->is_number) { /* 12A, 12B, ... */
            if (n > 5) return WARN_MSG("select: %s: '%s' invalid (string too long), will be ignored", selector_str[sel], val);
            if (n == 1) ++warn;
            if (!warn && (toupper((unsigned char)val[n - 1]) < 'A' || toupper((unsigned char)val[n - 1]) > 'Z')) ++warn;
            for (i = 0; !warn && i < n - 1; ++i)
                if (val[i] < '0' || val[i] > '9') ++warn;
            if (warn)
                return WARN_MSG("select: %s: '%s' invalid, should either be number (1, 2, 3) or number with insertion code (1A, 1B, ...), "
                                "will be ignored", selector_str[sel], val);
        }
        break;
    case S_CHAIN:
        if (n > 1) return WARN_MSG("select: %s: '%s' invalid (string too long), will be ignored", selector_str[sel], val);
        break;
    }
    return FREESASA_SUCCESS;
}

/* select_range(), src/selection.c:449-503; left/right NULL for the open ends */
static int select_range(struct parse *ps, enum selector sel, const struct ident *left, const struct ident *right, unsigned char *mask)
{
    const freesasa_structure *s = ps->s;
    int lower, upper, i;
    if (sel == S_RESI) {
        if ((left && !left->is_number) || (right && !right->is_number))
            return WARN_MSG("select: %s: range '%s-%s' invalid, needs to be two numbers, will be ignored", selector_str[sel],
                            left ? left->value : "", right ? right->value : "");
    } else {
        if (left->is_number != right->is_number || (!left->is_number && (strlen(left->value) > 1 || strlen(right->value) > 1)))
            return WARN_MSG("select: %s: range '%s-%s' invalid, should be two letters (A-C) or numbers (1-5), will be ignored",
                            selector_str[sel], left->value, right->value);
    }
    if (!left) {
        lower = atoi(s->label[0].res_number);
        upper = atoi(right->value);
    } else if (!right) {
        lower = atoi(left->value);
        upper = atoi(s->label[ps->n - 1].res_number);
    } else if (left->is_number) {
        lower = atoi(left->value);
        upper = atoi(right->value);
    } else {
        lower = (int)left->value[0];
        upper = (int)right->value[0];
    }
    for (i = 0; i < ps->n; ++i) {
        const int j = sel == S_RESI ? atoi(s->label[i].res_number) : (int)s->label[i].chain[0];
        if (j >= lower && j <= upper) mask[i] = 1;
    }
    return FREESASA_SUCCESS;
}

/* id : NUMBER | ID | '\-' NUMBER ; values are upper-cased (freesasa_selection_atom(), src/selection.c:104-137) */
static int parse_id(struct parse *ps, struct ident *id)
{
    size_t i;
    if (ps->sc.tok == T_MINUS) {
        next_token(&ps->sc);
        if (ps->sc.tok != T_NUMBER) return 0;
        id->is_number = 1;
        snprintf(id->value, sizeof id->value, "-%.250s", ps->sc.text);
    } else if (ps->sc.tok == T_NUMBER || ps->sc.tok == T_ID) {
        id->is_number = ps->sc.tok == T_NUMBER;
        snprintf(id->value, sizeof id->value, "%s", ps->sc.text);
    } else {
        return 0;
    }
    for (i = 0; id->value[i]; ++i) id->value[i] = (char)toupper((unsigned char)id->value[i]);
    next_token(&ps->sc);
    return 1;
}

static void apply_id(struct parse *ps, enum selector sel, const struct ident *id, unsigned char *mask)
{
    if (ps->dry) return;
    if (valid_id(sel, id) == FREESASA_SUCCESS) {
        select_id(ps, sel, id->value, mask);
    } else {
        WARN_MSG("select: %s: '%s' invalid %s", selector_str[sel], id->value, id->is_number ? "<number>" : "<id>");
        ++ps->warn;
    }
}

/* The argument of a selector: items joined by '+'.  resn/symbol/name take plain identifiers, resi and chain also ranges
 * (and resi open ranges).  Returns 0 on a syntax error. */
static int parse_items(struct parse *ps, enum selector sel, unsigned char *mask)
{
    const int ranges = sel == S_RESI || sel == S_CHAIN;
    for (;;) {
        struct ident a, b;
        if (ps->sc.tok == T_DASH) { /* '-' id : open on the left (resi only) */
            if (sel != S_RESI) return 0;
            next_token(&ps->sc);
            if (!parse_id(ps, &b)) return 0;
            if (!ps->dry && select_range(ps, sel, NULL, &b, mask) == FREESASA_WARN) ++ps->warn;
        } else {
            if (!parse_id(ps, &a)) return 0;
            if (ranges && ps->sc.tok == T_DASH) {
                next_token(&ps->sc);
                if (ps->sc.tok == T_MINUS || ps->sc.tok == T_NUMBER || ps->sc.tok == T_ID) {
                    if (!parse_id(ps, &b)) return 0;
                    if (!ps->dry && select_range(ps, sel, &a, &b, mask) == FREESASA_WARN) ++ps->warn;
                } else { /* id '-' : open on the right (resi only) */
                    if (sel != S_RESI) return 0;
                    if (!ps->dry && select_range(ps, sel, &a, NULL, mask) == FREESASA_WARN) ++ps->warn;
                }
            } else {
                apply_id(ps, sel, &a, mask);
            }
        }
        if (ps->sc.tok != T_PLUS) return 1;
        next_token(&ps->sc);
    }
}

static int parse_or(struct parse *ps, unsigned char *mask);

/* not-level and primaries */
#define MAX_NESTING 200
static int parse_unary(struct parse *ps, unsigned char *mask)
{
    int i, ok;
    enum selector sel;
    switch (ps->sc.tok) {
    case T_NOT:
        if (++ps->depth > MAX_NESTING) return 0;
        next_token(&ps->sc);
        ok = parse_unary(ps, mask);
        --ps->depth;
        if (!ok) return 0;
        for (i = 0; i < ps->n; ++i) mask[i] = !mask[i];
        return 1;
    case T_LPAR:
        if (++ps->depth > MAX_NESTING) return 0;
        next_token(&ps->sc);
        ok = parse_or(ps, mask);
        --ps->depth;
        if (!ok) return 0;
        if (ps->sc.tok != T_RPAR) return 0;
        next_token(&ps->sc);
        return 1;
    case T_RESN: sel = S_RESN; break;
    case T_RESI: sel = S_RESI; break;
    case T_SYMBOL: sel = S_SYMBOL; break;
    case T_NAME: sel = S_NAME; break;
    case T_CHAIN: sel = S_CHAIN; break;
    default: return 0;
    }
    next_token(&ps->sc);
    memset(mask, 0, (size_t)ps->n);
    return parse_items(ps, sel, mask);
}

static int parse_binary(struct parse *ps, unsigned char *mask, int is_or)
{
    unsigned char *rhs;
    int i;
    if (!(is_or ? parse_binary(ps, mask, 0) : parse_unary(ps, mask))) return 0;
    while (ps->sc.tok == (is_or ? T_OR : T_AND)) {
        next_token(&ps->sc);
        if (!(rhs = calloc((size_t)(ps->n > 0 ? ps->n : 1), 1))) {
            MEM_FAIL();
            ps->err = 1;
            return 0;
        }
        if (!(is_or ? parse_binary(ps, rhs, 0) : parse_unary(ps, rhs))) {
            free(rhs);
            return 0;
        }
        for (i = 0; i < ps->n; ++i) mask[i] = is_or ? (mask[i] || rhs[i]) : (mask[i] && rhs[i]);
        free(rhs);
    }
    return 1;
}
static int parse_or(struct parse *ps, unsigned char *mask) { return parse_binary(ps, mask, 1); }

/* select_area_impl(), src/selection.c:679-749: number of atoms looked at, FREESASA_WARN or FREESASA_FAIL */
static int select_area(const char *command, char *name, double *area, const freesasa_structure *structure, const freesasa_result *result)
{
    struct parse ps;
    unsigned char *mask;
    double sasa = 0;
    int ok, j;

    assert(name);
    assert(area);
    assert(command);
    assert(structure);
    assert(result);
    assert(freesasa_structure_n(structure) == result->n_atoms);
    *area = 0;
    name[0] = '\0';
    memset(&ps, 0, sizeof ps);
    ps.s = structure;
    ps.n = result->n_atoms;
    if (!(mask = calloc((size_t)(ps.n > 0 ? ps.n : 1), 1))) return FAIL_MSG("%s", "");

    for (ps.dry = 1, ok = 1; ok && ps.dry >= 0; --ps.dry) { /* first the syntax, then the evaluation */
        ps.sc.p = command;
        next_token(&ps.sc);
        ok = ps.sc.tok == T_SELID;
        if (ok) {
            snprintf(name, MAX_SELECTION_NAME + 1, "%.50s", ps.sc.text);
            next_token(&ps.sc);
            ok = ps.sc.tok == T_COMMA;
        }
        if (ok) {
            next_token(&ps.sc);
            ok = parse_or(&ps, mask) && ps.sc.tok == T_END;
        }
    }
    if (!ok) {
        free(mask);
        name[0] = '\0';
        if (!ps.err) fsb_report(FREESASA_FAIL, NULL, 0, "syntax error");
        return FAIL_MSG("problems parsing expression '%s'", command);
    }
    for (j = 0; j < ps.n; ++j) sasa += mask[j] * result->sasa[j];
    *area = sasa;
    free(mask);
    if (ps.warn) return WARN_MSG("in %s(): There were warnings", "select_area_impl");
    return ps.n;
}

/* ---- public API (src/freesasa.h:610-690, src/selection.c:751-880) ------------------------------------------------ */
static freesasa_selection *selection_alloc(const char *name, const char *command)
{
    freesasa_selection *sel = calloc(1, sizeof *sel);
    if (sel == NULL || !(sel->name = strdup(name)) || !(sel->command = strdup(command))) {
        MEM_FAIL();
        freesasa_selection_free(sel);
        return NULL;
    }
    return sel;
}

void freesasa_selection_free(freesasa_selection *selection)
{
    if (selection != NULL) {
        free(selection->name);
        free(selection->command);
        free(selection);
    }
}

freesasa_selection *freesasa_selection_clone(const freesasa_selection *src)
{
    freesasa_selection *cpy = selection_alloc(src->name, src->command);
    if (cpy == NULL) {
        FAIL_MSG("%s", "");
        return NULL;
    }
    cpy->area = src->area;
    cpy->n_atoms = src->n_atoms;
    return cpy;
}

const char *freesasa_selection_name(const freesasa_selection *selection) { return selection->name; }
const char *freesasa_selection_command(const freesasa_selection *selection) { return selection->command; }
double freesasa_selection_area(const freesasa_selection *selection) { return selection->area; }
int freesasa_selection_n_atoms(const freesasa_selection *selection) { return selection->n_atoms; }

freesasa_selection *freesasa_selection_new(const char *command, const freesasa_structure *structure, const freesasa_result *result)
{
    char name[MAX_SELECTION_NAME + 1];
    double area;
    freesasa_selection *selection;
    const int n_atoms = select_area(command, name, &area, structure, result);
    if (n_atoms == FREESASA_FAIL) {
        FAIL_MSG("%s", "");
        return NULL;
    }
    if (!(selection = selection_alloc(name, command))) return NULL;
    selection->area = area;
    selection->n_atoms = n_atoms;
    return selection;
}

int freesasa_select_area(const char *command, char *name, double *area, const freesasa_structure *structure,
                         const freesasa_result *result)
{
    const int ret = select_area(command, name, area, structure, result);
    return ret >= 0 ? FREESASA_SUCCESS : ret;
}
