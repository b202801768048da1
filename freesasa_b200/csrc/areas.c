/* freesasa_b200/csrc/areas.c — per-atom SASA -> areas of residues, chains, structures (the result tree).
 *
 * Scope row f-3 of SURVEY.md §8(f).  Mirrors the reference's node API (src/freesasa.h:1449-1850, src/node.c) and its
 * class sums (freesasa_result_classes(), src/classifier.c:829-838): same node types, same parent/children/next
 * topology, same names and properties, and the same areas TO THE BIT — every sum is formed in the reference's order
 * (atoms of a residue in sequence, residues of a chain in sequence, chains of a structure in sequence; src/node.c:148-
 * 176,718-777), which is what makes a sum of doubles reproducible.
 *
 * What differs is the construction.  The reference builds the tree with one malloc per node, one per area and four to
 * five strdup per atom (name, chain, residue number, residue name, PDB line; src/node.c:214-277) and frees it node by
 * node.  Here one result (freesasa_tree_add_result) is ONE allocation: a header, the node array (result, structure,
 * chains, residues, atoms — in that order, children contiguous), the area array and a string pool in which the atoms of
 * a residue share the residue's strings.  PDB lines are not copied: the block keeps a reference on the text the
 * structure was read from and materialises NUL-terminated lines the first time one is asked for.  Building the tree for
 * 100k atoms is a single pass of ~25 ns/atom; freeing it is one free().
 *
 * The segmented sums stay on the host on purpose: with host-resident results (the drop-in contract returns `sasa[]` in
 * malloc'd host memory) a device epilogue would have to upload 5 B/atom of keys to save ~1 ns/atom of
 * additions.
 */
#include "host_internal.h"

#include <assert.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

struct tree_block;

struct freesasa_node {
    const char *name;
    freesasa_nodetype type;
    freesasa_nodearea *area;
    freesasa_node *parent, *children, *next;
    union {
        struct {
            int is_polar, is_bb, index; /* index: atom number inside its block */
            double radius;
            const char *chain, *res_number, *res_name;
        } atom;
        struct {
            int n_atoms;
            const char *number;
            const freesasa_nodearea *reference;
        } residue;
        struct {
            int n_residues;
        } chain;
        struct {
            int n_chains, n_atoms, model;
            const char *chain_labels;
            freesasa_result *result;
            freesasa_selection **selection; /* NULL-terminated array of clones, or NULL (src/node.c:32-40) */
        } structure;
        struct {
            const char *classified_by;
            freesasa_parameters parameters;
            int n_structures;
            struct tree_block *block;
        } result;
    } p;
};

/* one freesasa_tree_add_result(): everything below lives in the same allocation as this header */
struct tree_block {
    int n_atoms;
    struct shared_text *text; /* referenced, may be NULL */
    long *line_at;
    unsigned char *line_len;
    char *volatile lines; /* separately allocated on demand */
    freesasa_result result;
};

const freesasa_nodearea freesasa_nodearea_null = {NULL, 0, 0, 0, 0, 0, 0};

/* ---- area arithmetic (src/node.c:718-777) --------------------------------------------------------------------- */
static inline void atom_area(freesasa_nodearea *area, double a, int is_bb, int cls)
{
    *area = freesasa_nodearea_null;
    area->total = a;
    if (is_bb)
        area->main_chain = a;
    else
        area->side_chain = a;
    switch (cls) {
    case FREESASA_ATOM_APOLAR: area->apolar = a; break;
    case FREESASA_ATOM_POLAR: area->polar = a; break;
    case FREESASA_ATOM_UNKNOWN: area->unknown = a; break;
    }
}

int freesasa_atom_nodearea(freesasa_nodearea *area, const freesasa_structure *structure, const freesasa_result *result,
                           int atom_index)
{
    atom_area(area, result->sasa[atom_index], FSB_CLS_BACKBONE(structure->cls[atom_index]), FSB_CLS_CLASS(structure->cls[atom_index]));
    return FREESASA_SUCCESS;
}

void freesasa_add_nodearea(freesasa_nodearea *sum, const freesasa_nodearea *term)
{
    sum->total += term->total;
    sum->side_chain += term->side_chain;
    sum->main_chain += term->main_chain;
    sum->polar += term->polar;
    sum->apolar += term->apolar;
    sum->unknown += term->unknown;
}

void freesasa_range_nodearea(freesasa_nodearea *area, const freesasa_structure *structure, const freesasa_result *result,
                             int first_atom, int last_atom)
{
    freesasa_nodearea term;
    int i;
    assert(area);
    assert(structure);
    assert(result);
    assert(first_atom <= last_atom);
    for (i = first_atom; i <= last_atom; ++i) {
        freesasa_atom_nodearea(&term, structure, result, i);
        freesasa_add_nodearea(area, &term);
    }
}

/* src/classifier.c:829-838 */
freesasa_nodearea freesasa_result_classes(const freesasa_structure *structure, const freesasa_result *result)
{
    freesasa_nodearea area = {"whole-structure", 0, 0, 0, 0, 0, 0};
    freesasa_range_nodearea(&area, structure, result, 0, freesasa_structure_n(structure) - 1);
    return area;
}

/* ---- construction ------------------------------------------------------------------------------------------- */
static size_t align8(size_t n) { return (n + 7) & ~(size_t)7; }

freesasa_node *freesasa_tree_new(void)
{
    freesasa_node *root = calloc(1, sizeof *root);
    if (root == NULL) {
        MEM_FAIL();
        return NULL;
    }
    root->type = FREESASA_NODE_ROOT;
    return root;
}

static void block_free(freesasa_node *result_node)
{
    struct tree_block *b = result_node->p.result.block;
    freesasa_selection **sel = result_node->children->p.structure.selection;
    if (sel) {
        freesasa_selection **it;
        for (it = sel; *it; ++it) freesasa_selection_free(*it);
        free(sel);
    }
    fsb_text_release(b->text);
    free(b->lines);
    free(b); /* nodes, areas, strings and the cloned result all live in this allocation */
}

/* Everything the per-residue-range workers of freesasa_tree_add_result() need. */
struct build {
    const freesasa_structure *s;
    const freesasa_result *result;
    struct tree_block *b;
    freesasa_node *res_nodes, *atom_nodes;
    freesasa_nodearea *res_areas, *atom_areas, *refs;
    char *pool_res, *pool_atom, *pool_alt; /* 14 B per residue (name 4, number 6, own chain label 4); 5 B and 4 B per atom */
};

/* Residues [n_res*part/n_parts, n_res*(part+1)/n_parts) and their atoms: nodes, areas, strings, and the slices of the
 * per-atom tables the block keeps.  Touches only its own slice of every array (first touch in parallel). */
static void build_part(int part, int n_parts, void *arg)
{
    const struct build *q = arg;
    const freesasa_structure *s = q->s;
    const int n = s->n, n_res = s->n_res;
    const int r0 = (int)((long long)n_res * part / n_parts), r1 = (int)((long long)n_res * (part + 1) / n_parts);
    const int a0 = r0 < n_res ? s->res_first[r0] : n, a1 = r1 < n_res ? s->res_first[r1] : n;
    int r, i;

    if (r0 >= r1) return;
    memcpy(q->b->line_at + a0, s->line_at + a0, (size_t)(a1 - a0) * sizeof(long));
    memcpy(q->b->line_len + a0, s->line_len + a0, (size_t)(a1 - a0));
    memcpy(q->b->result.sasa + a0, q->result->sasa + a0, (size_t)(a1 - a0) * sizeof(double)); /* freesasa_result_clone() */
    for (i = a0; i < a1; ++i) {
        freesasa_node *a = &q->atom_nodes[i];
        const struct atom_label *l = &s->label[i];
        char *name = q->pool_atom + 5 * (size_t)i;
        memcpy(name, l->name, 5);
        a->name = name;
        a->type = FREESASA_NODE_ATOM;
        a->area = &q->atom_areas[i];
        a->parent = &q->res_nodes[s->res_index[i]];
        a->children = NULL;
        a->next = a + 1; /* the last atom of each residue is cut below */
        a->p.atom.is_polar = FSB_CLS_CLASS(s->cls[i]) == FREESASA_ATOM_POLAR;
        a->p.atom.is_bb = FSB_CLS_BACKBONE(s->cls[i]); /* decided once per distinct name while reading */
        a->p.atom.index = i;
        a->p.atom.radius = s->radius[i];
        atom_area(a->area, q->result->sasa[i], a->p.atom.is_bb, FSB_CLS_CLASS(s->cls[i])); /* name stays NULL, as in src/node.c:269-270 */
    }
    for (r = r0; r < r1; ++r) {
        freesasa_node *res = &q->res_nodes[r];
        const int first = s->res_first[r], last = r == n_res - 1 ? n - 1 : s->res_first[r + 1] - 1;
        const struct atom_label *l = &s->label[first];
        char *str = q->pool_res + 14 * (size_t)r;
        const char *chain = res->parent->name; /* set by the serial chain pass */
        freesasa_nodearea *sum = &q->res_areas[r];
        res->type = FREESASA_NODE_RESIDUE;
        memcpy(str, l->res_name, 4);
        memcpy(str + 4, l->res_number, 6);
        res->name = str;
        /* an atom reports its own chain label; it differs from the chain node it hangs under only when a chain label
         * comes back after another one (A, B, A: the second run of A is filed under B, src/structure.c:1303-1325) */
        if (memcmp(l->chain, chain, 4) != 0) {
            memcpy(str + 10, l->chain, 4);
            chain = str + 10;
        }
        res->p.residue.number = str + 4;
        res->p.residue.n_atoms = last - first + 1;
        res->p.residue.reference = NULL;
        if (s->res_has_ref[r]) { /* a copy: the tree may outlive the structure (src/node.c:303-315) */
            q->refs[r] = s->res_ref[r];
            res->p.residue.reference = &q->refs[r];
        }
        res->children = &q->atom_nodes[first];
        res->area = sum;
        *sum = freesasa_nodearea_null;
        sum->name = res->name;
        for (i = first; i <= last; ++i) {
            freesasa_node *a = &q->atom_nodes[i];
            freesasa_add_nodearea(sum, a->area);
            a->p.atom.res_name = res->name; /* residues are delimited by number and chain only: names may differ inside */
            if (memcmp(s->label[i].res_name, l->res_name, 4) != 0) {
                char *alt = q->pool_alt + 4 * (size_t)i;
                memcpy(alt, s->label[i].res_name, 4);
                a->p.atom.res_name = alt;
            }
            a->p.atom.res_number = str + 4;
            a->p.atom.chain = chain;
        }
        q->atom_nodes[last].next = NULL;
    }
}

/* freesasa_tree_add_result(), src/node.c:441-476, with node_structure/node_chain/node_residue/node_atom
 * (src/node.c:214-409) fused: a serial pass over the chains, one pass over residues and atoms (split over threads for
 * large structures: every string lives at an index-derived place in the pool, so the parts are independent), then the
 * chain and structure sums in order */
static int tree_add_result(freesasa_node *tree, const freesasa_result *result, const freesasa_structure *s, const char *name,
                           int allow_threads)
{
    const int n = s->n, n_res = s->n_res, n_chains = s->n_chains;
    const size_t n_nodes = 2 + (size_t)n_chains + (size_t)n_res + (size_t)n;
    const size_t name_len = name ? strlen(name) + 1 : 0, cls_len = strlen(s->classifier_name ? s->classifier_name : "") + 1;
    const size_t labels_len = strlen(s->short_labels ? s->short_labels : "") + 1;
    /* string pool: result name, classifier name, chain labels (twice: structure name and property), 4 per chain; then
     * the index-addressed regions: 14 per residue, 5 + 4 per atom (the 4: a residue name that differs from the residue
     * node's — untouched address space otherwise) */
    const size_t fixed = name_len + cls_len + 2 * labels_len + 4 * (size_t)n_chains;
    const size_t pool = fixed + 14 * (size_t)n_res + 9 * (size_t)n;
    size_t off_nodes, off_areas, off_refs, off_sasa, off_lines, off_len, off_pool, total;
    char *mem, *str;
    struct tree_block *b;
    struct build q;
    freesasa_node *nodes, *rnode, *snode, *chain_nodes, *res_nodes, *atom_nodes;
    freesasa_nodearea *areas;
    int c, r, parts;

    assert(tree);
    assert(tree->type == FREESASA_NODE_ROOT);
    if (s->classifier_name == NULL || n == 0) return FAIL_MSG("structure without atoms");

    off_nodes = align8(sizeof(struct tree_block));
    off_areas = off_nodes + n_nodes * sizeof(freesasa_node);
    off_refs = off_areas + n_nodes * sizeof(freesasa_nodearea);
    off_sasa = off_refs + (size_t)n_res * sizeof(freesasa_nodearea);
    off_lines = off_sasa + (size_t)n * sizeof(double);
    off_len = off_lines + (size_t)n * sizeof(long);
    off_pool = align8(off_len + (size_t)n);
    total = off_pool + pool;
    if (!(mem = malloc(total))) {
        MEM_FAIL();
        return FAIL_MSG("%s", "");
    }
    b = (struct tree_block *)mem;
    nodes = (freesasa_node *)(mem + off_nodes);
    areas = (freesasa_nodearea *)(mem + off_areas);
    str = mem + off_pool;

    b->n_atoms = n;
    b->text = s->text;
    if (b->text) __atomic_add_fetch(&b->text->refs, 1, __ATOMIC_RELAXED);
    b->line_at = (long *)(mem + off_lines);
    b->line_len = (unsigned char *)(mem + off_len);
    b->lines = NULL;
    b->result = *result;
    b->result.sasa = (double *)(mem + off_sasa);

    rnode = nodes;
    snode = nodes + 1;
    chain_nodes = nodes + 2;
    res_nodes = chain_nodes + n_chains;
    atom_nodes = res_nodes + n_res;

#define POOL(dst, src, len)            \
    do {                               \
        memcpy(str, (src), (len));     \
        (dst) = str;                   \
        str += (len);                  \
    } while (0)

    /* result node: no area, no parent (src/node.c:148-156,441-470) */
    memset(rnode, 0, sizeof *rnode);
    rnode->type = FREESASA_NODE_RESULT;
    if (name) POOL(rnode->name, name, name_len);
    POOL(rnode->p.result.classified_by, s->classifier_name, cls_len);
    rnode->p.result.parameters = result->parameters;
    rnode->p.result.n_structures = 1;
    rnode->p.result.block = b;
    rnode->children = snode;

    memset(snode, 0, sizeof *snode);
    snode->type = FREESASA_NODE_STRUCTURE;
    POOL(snode->name, s->short_labels, labels_len);
    POOL(snode->p.structure.chain_labels, s->short_labels, labels_len);
    snode->p.structure.n_chains = n_chains;
    snode->p.structure.n_atoms = n;
    snode->p.structure.model = s->model;
    snode->p.structure.result = &b->result;
    snode->parent = rnode;
    snode->children = chain_nodes;
    snode->area = &areas[1];

    /* chains: names and residue ranges first (the residues need their parent), sums afterwards */
    for (c = 0; c < n_chains; ++c) {
        freesasa_node *ch = &chain_nodes[c];
        const int first_atom = s->chain_first[c], last_atom = c == n_chains - 1 ? n - 1 : s->chain_first[c + 1] - 1;
        const int first_res = s->res_index[first_atom], last_res = s->res_index[last_atom];
        ch->type = FREESASA_NODE_CHAIN;
        POOL(ch->name, s->chain_label[c], 4);
        ch->p.chain.n_residues = last_res - first_res + 1;
        ch->parent = snode;
        ch->children = &res_nodes[first_res];
        ch->next = c == n_chains - 1 ? NULL : ch + 1;
        ch->area = &areas[2 + c];
        for (r = first_res; r <= last_res; ++r) {
            res_nodes[r].parent = ch;
            res_nodes[r].next = r == last_res ? NULL : &res_nodes[r + 1];
        }
    }
#undef POOL
    assert((size_t)(str - (mem + off_pool)) <= fixed);

    q.s = s;
    q.result = result;
    q.b = b;
    q.res_nodes = res_nodes;
    q.atom_nodes = atom_nodes;
    q.res_areas = &areas[2 + n_chains];
    q.atom_areas = &areas[2 + n_chains + n_res];
    q.refs = (freesasa_nodearea *)(mem + off_refs);
    q.pool_res = mem + off_pool + fixed;
    q.pool_atom = q.pool_res + 14 * (size_t)n_res;
    q.pool_alt = q.pool_atom + 5 * (size_t)n;
    {
        /* FREESASA_B200_PARALLEL_MIN_ATOMS: test hook, lets the test-suite drive small structures through the threaded build */
        const char *env = getenv("FREESASA_B200_PARALLEL_MIN_ATOMS");
        const int min_atoms = env ? atoi(env) : 20000, per_part = min_atoms / 2 > 16 ? min_atoms / 2 : 16;
        parts = allow_threads && n >= min_atoms ? fsb_hardware_threads() : 1;
        if (parts > n / per_part) parts = n / per_part > 0 ? n / per_part : 1;
    }
    if (parts > 1)
        fsb_parallel_run(parts, build_part, &q);
    else
        build_part(0, 1, &q);

    for (c = 0; c < n_chains; ++c) {
        freesasa_node *ch = &chain_nodes[c], *res;
        *ch->area = freesasa_nodearea_null;
        ch->area->name = ch->name;
        for (res = ch->children; res; res = res->next) freesasa_add_nodearea(ch->area, res->area);
    }
    *snode->area = freesasa_nodearea_null;
    snode->area->name = snode->name;
    for (c = 0; c < n_chains; ++c) freesasa_add_nodearea(snode->area, chain_nodes[c].area);

    /* prepend to the root's list (src/node.c:467-468) */
    rnode->next = tree->children;
    tree->children = rnode;
    return FREESASA_SUCCESS;
}

int freesasa_tree_add_result(freesasa_node *tree, const freesasa_result *result, const freesasa_structure *s, const char *name)
{
    return tree_add_result(tree, result, s, name, 1);
}

freesasa_node *freesasa_tree_init(const freesasa_result *result, const freesasa_structure *structure, const char *name)
{
    freesasa_node *tree = freesasa_tree_new();
    if (tree == NULL) {
        FAIL_MSG("%s", "");
    } else if (freesasa_tree_add_result(tree, result, structure, name) == FREESASA_FAIL) {
        FAIL_MSG("%s", "");
        freesasa_node_free(tree);
        tree = NULL;
    }
    return tree;
}

/* src/node.c:478-503 */
int freesasa_tree_join(freesasa_node *tree1, freesasa_node **tree2)
{
    freesasa_node *child;
    assert(tree1);
    assert(tree2);
    assert(*tree2);
    assert(tree1->type == FREESASA_NODE_ROOT);
    assert((*tree2)->type == FREESASA_NODE_ROOT);
    child = tree1->children;
    if (child != NULL) {
        while (child->next) child = child->next;
        child->next = (*tree2)->children;
    } else {
        tree1->children = (*tree2)->children;
    }
    free(*tree2);
    *tree2 = NULL;
    return FREESASA_SUCCESS;
}

/* src/node.c:505-513.  Result nodes have no parent in the reference either (src/node.c:441-470), so a result taken
 * out of context can be freed on its own there; here as well. */
int freesasa_node_free(freesasa_node *root)
{
    freesasa_node *r, *next;
    if (root == NULL) return FREESASA_SUCCESS;
    if (root->parent || (root->type != FREESASA_NODE_ROOT && root->type != FREESASA_NODE_RESULT))
        return FAIL_MSG("can't free node that isn't the root of its tree");
    if (root->type == FREESASA_NODE_RESULT) {
        block_free(root);
        return FREESASA_SUCCESS;
    }
    for (r = root->children; r; r = next) {
        next = r->next;
        block_free(r);
    }
    free(root);
    return FREESASA_SUCCESS;
}

/* src/freesasa.c:155-182 */
freesasa_node *freesasa_calc_tree(const freesasa_structure *structure, const freesasa_parameters *parameters, const char *name)
{
    freesasa_node *tree = NULL;
    freesasa_result *result;
    assert(structure);
    result = freesasa_calc(freesasa_structure_xyz(structure), freesasa_structure_radius(structure), parameters);
    if (result != NULL)
        tree = freesasa_tree_init(result, structure, name);
    else
        FAIL_MSG("%s", "");
    if (tree == NULL) FAIL_MSG("%s", "");
    freesasa_result_free(result);
    return tree;
}

/* Additive (row f-2): what the CLI's loop over structures (src/main.cc:334-362: freesasa_calc_tree per model / chain
 * group) becomes with a batched engine — every structure of an array in ONE device pass, then one tree per structure, the
 * trees built concurrently on the host's cores.  trees[k] receives a tree as freesasa_calc_tree(structures[k], parameters,
 * names ? names[k] : NULL) would return it; on failure no tree is left allocated.  Messages come out in structure order. */
struct tree_batch {
    int n, next;
    freesasa_structure *const *structures;
    freesasa_result **results;
    const char *const *names;
    freesasa_node **trees;
    struct fsb_capture *messages;
};
static void tree_batch_worker(int part, int n_parts, void *arg)
{
    struct tree_batch *w = arg;
    (void)part;
    (void)n_parts;
    for (;;) {
        const int k = __atomic_fetch_add(&w->next, 1, __ATOMIC_RELAXED);
        freesasa_node *tree;
        if (k >= w->n) break;
        fsb_capture_current = &w->messages[k];
        tree = freesasa_tree_new();
        if (tree && tree_add_result(tree, w->results[k], w->structures[k], w->names ? w->names[k] : NULL, 0) == FREESASA_FAIL) {
            freesasa_node_free(tree);
            tree = NULL;
        }
        w->trees[k] = tree;
        fsb_capture_current = NULL;
    }
}
/* One tree per (structure, result) pair, built concurrently; messages replayed in order.  On failure no tree is left. */
int fsb_trees_from_results(int n_struct, freesasa_structure *const *structures, freesasa_result *const *results,
                           const char *const *names, freesasa_node **trees)
{
    struct tree_batch w;
    int k, failed = 0, threads;
    for (k = 0; k < n_struct; ++k) trees[k] = NULL;
    memset(&w, 0, sizeof w);
    w.n = n_struct;
    w.structures = structures;
    w.results = (freesasa_result **)results;
    w.names = names;
    w.trees = trees;
    if (!(w.messages = calloc((size_t)n_struct, sizeof *w.messages))) return MEM_FAIL();
    threads = fsb_hardware_threads();
    fsb_parallel_run(threads < n_struct ? threads : n_struct, tree_batch_worker, &w);
    for (k = 0; k < n_struct; ++k) {
        fsb_capture_flush(&w.messages[k]);
        if (trees[k] == NULL) failed = 1;
    }
    free(w.messages);
    if (failed) {
        for (k = 0; k < n_struct; ++k) {
            freesasa_node_free(trees[k]);
            trees[k] = NULL;
        }
        return FAIL_MSG("%s", "");
    }
    return FREESASA_SUCCESS;
}

int freesasa_calc_tree_batch(int n_struct, freesasa_structure *const *structures, const freesasa_parameters *parameters,
                             const char *const *names, freesasa_node **trees)
{
    freesasa_result **results;
    int k, rc;
    if (n_struct <= 0 || !structures || !trees) return FAIL_MSG("invalid batch arguments");
    for (k = 0; k < n_struct; ++k) trees[k] = NULL;
    if (!(results = calloc((size_t)n_struct, sizeof *results))) return MEM_FAIL();
    if (freesasa_calc_structure_batch(n_struct, structures, parameters, results) != FREESASA_SUCCESS) {
        free(results);
        return FAIL_MSG("%s", "");
    }
    rc = fsb_trees_from_results(n_struct, structures, results, names, trees);
    for (k = 0; k < n_struct; ++k) freesasa_result_free(results[k]);
    free(results);
    return rc;
}

/* ---- accessors (src/node.c:515-716) ----------------------------------------------------------------------- */
const freesasa_nodearea *freesasa_node_area(const freesasa_node *node)
{
    assert(node->type != FREESASA_NODE_ROOT);
    return node->area;
}
freesasa_node *freesasa_node_children(freesasa_node *node) { return node->children; }
freesasa_node *freesasa_node_next(freesasa_node *node) { return node->next; }
freesasa_node *freesasa_node_parent(freesasa_node *node) { return node->parent; }
freesasa_nodetype freesasa_node_type(const freesasa_node *node) { return node->type; }
const char *freesasa_node_name(const freesasa_node *node) { return node->name; }
const char *freesasa_node_classified_by(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_RESULT);
    return node->p.result.classified_by;
}
int freesasa_node_atom_is_polar(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_ATOM);
    return node->p.atom.is_polar;
}
int freesasa_node_atom_is_mainchain(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_ATOM);
    return node->p.atom.is_bb;
}
double freesasa_node_atom_radius(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_ATOM);
    return node->p.atom.radius;
}
const char *freesasa_node_atom_residue_number(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_ATOM);
    return node->p.atom.res_number;
}
const char *freesasa_node_atom_residue_name(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_ATOM);
    return node->p.atom.res_name;
}
const char *freesasa_node_atom_chain(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_ATOM);
    return node->p.atom.chain;
}
/* the PDB record the atom was read from (with its newline), NULL for atoms added by hand; copies are made for the
 * whole block the first time any line is requested */
const char *freesasa_node_atom_pdb_line(const freesasa_node *node)
{
    static pthread_mutex_t once = PTHREAD_MUTEX_INITIALIZER;
    const freesasa_node *r;
    struct tree_block *b;
    int i;
    assert(node->type == FREESASA_NODE_ATOM);
    r = node->parent->parent->parent->parent; /* residue, chain, structure, result */
    b = r->p.result.block;
    i = node->p.atom.index;
    if (b->line_at[i] < 0) return NULL;
    if (__atomic_load_n(&b->lines, __ATOMIC_ACQUIRE) == NULL) {
        pthread_mutex_lock(&once);
        if (b->lines == NULL) {
            char *all = malloc((size_t)b->n_atoms * LINE_MAX_STRL);
            int k;
            if (all == NULL) {
                pthread_mutex_unlock(&once);
                MEM_FAIL();
                return NULL;
            }
            for (k = 0; k < b->n_atoms; ++k) {
                const int len = b->line_at[k] >= 0 ? b->line_len[k] : 0;
                if (len) memcpy(all + (size_t)k * LINE_MAX_STRL, b->text->data + b->line_at[k], (size_t)len);
                all[(size_t)k * LINE_MAX_STRL + len] = '\0';
            }
            __atomic_store_n((char **)&b->lines, all, __ATOMIC_RELEASE);
        }
        pthread_mutex_unlock(&once);
    }
    return b->lines + (size_t)i * LINE_MAX_STRL;
}
int freesasa_node_residue_n_atoms(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_RESIDUE);
    return node->p.residue.n_atoms;
}
const char *freesasa_node_residue_number(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_RESIDUE);
    return node->p.residue.number;
}
const freesasa_nodearea *freesasa_node_residue_reference(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_RESIDUE);
    return node->p.residue.reference;
}
int freesasa_node_chain_n_residues(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_CHAIN);
    return node->p.chain.n_residues;
}
int freesasa_node_structure_n_chains(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_STRUCTURE);
    return node->p.structure.n_chains;
}
int freesasa_node_structure_n_atoms(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_STRUCTURE);
    return node->p.structure.n_atoms;
}
int freesasa_node_structure_model(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_STRUCTURE);
    return node->p.structure.model;
}
const char *freesasa_node_structure_chain_labels(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_STRUCTURE);
    return node->p.structure.chain_labels;
}
const freesasa_result *freesasa_node_structure_result(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_STRUCTURE);
    return node->p.structure.result;
}
/* src/node.c:670-709: the node keeps its own copy of every selection added */
int freesasa_node_structure_add_selection(freesasa_node *node, const freesasa_selection *selection)
{
    freesasa_selection **sel;
    int n = 0;
    assert(node->type == FREESASA_NODE_STRUCTURE);
    sel = node->p.structure.selection;
    if (sel)
        while (sel[n]) ++n;
    sel = realloc(sel, sizeof(freesasa_selection *) * (size_t)(n + 2));
    if (sel == NULL) return MEM_FAIL();
    node->p.structure.selection = sel;
    sel[n] = freesasa_selection_clone(selection);
    sel[n + 1] = NULL;
    if (sel[n] == NULL) return FAIL_MSG("%s", "");
    return FREESASA_SUCCESS;
}
const freesasa_selection **freesasa_node_structure_selections(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_STRUCTURE);
    return (const freesasa_selection **)node->p.structure.selection;
}
const freesasa_parameters *freesasa_node_result_parameters(const freesasa_node *node)
{
    assert(node->type == FREESASA_NODE_RESULT);
    return &node->p.result.parameters;
}

/* ---- per-atom writer (scope row f-4): PDB with radius and SASA in the occupancy / B-factor columns ------------------ */
const char *freesasa_string = "freesasa-b200 (FreeSASA 2.1.3 API)";

/* write_pdb_impl(), src/pdb.c:284-345: the first 54 columns of the original record, then "%6.2f%6.2f" of radius and
 * area (which ends the line: the reference's sprintf terminates the buffer there), a TER record numbered one past the
 * last atom's serial, ENDMDL */
static int write_structure_pdb(FILE *output, freesasa_node *structure)
{
    char buf[81], serial[6];
    freesasa_node *chain, *residue, *atom;
    const char *last_res_name = NULL, *last_res_number = NULL, *last_chain = NULL;
    const int model = freesasa_node_structure_model(structure);

    if (model > 0)
        fprintf(output, "MODEL     %4d\n", model);
    else
        fprintf(output, "MODEL        1\n");
    memset(buf, 0, sizeof buf);
    for (chain = structure->children; chain; chain = chain->next) {
        for (residue = chain->children; residue; residue = residue->next) {
            for (atom = residue->children; atom; atom = atom->next) {
                const char *line = freesasa_node_atom_pdb_line(atom);
                if (line == NULL) return FAIL_MSG("PDB input not valid or not present");
                strncpy(buf, line, 80);
                sprintf(&buf[54], "%6.2f%6.2f", atom->p.atom.radius, atom->area->total);
                fprintf(output, "%s\n", buf);
            }
            last_res_name = residue->name;
            last_res_number = residue->p.residue.number;
        }
        last_chain = chain->name;
    }
    memcpy(serial, &buf[6], 5);
    serial[5] = '\0';
    fprintf(output, "TER   %5d     %4s %c%5s\nENDMDL\n", atoi(serial) + 1, last_res_name, last_chain[0], last_res_number);
    fflush(output);
    if (ferror(output)) return FAIL_MSG("write error");
    return FREESASA_SUCCESS;
}

/* src/pdb.c:347-375 */
int freesasa_write_pdb(FILE *output, freesasa_node *root)
{
    freesasa_node *result, *structure;
    assert(output);
    assert(root);
    assert(root->type == FREESASA_NODE_ROOT);
    fprintf(output, "REMARK 999 This PDB file was generated by %s.\n", freesasa_string);
    fprintf(output, "REMARK 999 In the ATOM records temperature factors have been\n"
                    "REMARK 999 replaced by the SASA of the atom, and the occupancy\n"
                    "REMARK 999 by the radius used in the calculation.\n");
    for (result = root->children; result; result = result->next)
        for (structure = result->children; structure; structure = structure->next)
            if (write_structure_pdb(output, structure) == FREESASA_FAIL) return FAIL_MSG("%s", "");
    return FREESASA_SUCCESS;
}

/* ---- per-residue-type and per-residue listings (src/log.c:150-246; the CLI's --format=res / --format=seq) ------------- */
int freesasa_write_res(FILE *log, freesasa_node *root)
{
    const int n_types = freesasa_classify_n_residue_types() + 1;
    double *area = malloc(sizeof(double) * (size_t)n_types);
    freesasa_node *result, *structure, *chain, *residue;
    int i;
    assert(log);
    assert(root);
    assert(root->type == FREESASA_NODE_ROOT);
    if (area == NULL) return MEM_FAIL();
    for (result = root->children; result; result = result->next) {
        for (i = 0; i < n_types; ++i) area[i] = 0;
        for (structure = result->children; structure; structure = structure->next)
            for (chain = structure->children; chain; chain = chain->next)
                for (residue = chain->children; residue; residue = residue->next)
                    area[freesasa_classify_residue(residue->name)] += residue->area->total; /* in tree order, as the reference */
        fprintf(log, "# Residue types in %s\n", result->name);
        for (i = 0; i < n_types - 1; ++i)
            if (i < 20 || area[i] > 0) fprintf(log, "RES %s : %10.2f\n", freesasa_classify_residue_name(i), area[i]);
        fprintf(log, "\n");
    }
    free(area);
    fflush(log);
    if (ferror(log)) return FAIL_MSG("write error");
    return FREESASA_SUCCESS;
}

int freesasa_write_seq(FILE *log, freesasa_node *root)
{
    freesasa_node *result, *structure, *chain, *residue;
    assert(log);
    assert(root);
    assert(root->type == FREESASA_NODE_ROOT);
    for (result = root->children; result; result = result->next) {
        fprintf(log, "# Residues in %s\n", result->name);
        for (structure = result->children; structure; structure = structure->next)
            for (chain = structure->children; chain; chain = chain->next)
                for (residue = chain->children; residue; residue = residue->next)
                    fprintf(log, "SEQ %s %s %s : %7.2f\n", chain->name, residue->p.residue.number, residue->name, residue->area->total);
        fprintf(log, "\n");
    }
    fflush(log);
    if (ferror(log)) return FAIL_MSG("write error");
    return FREESASA_SUCCESS;
}
