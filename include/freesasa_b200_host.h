/* include/freesasa_b200_host.h — the hot-path subset of the FreeSASA C API, B200-backed.
 *
 * libfreesasa_b200_host.so exports, under the reference's own names and with ABI-identical
 * structs, exactly the functions that sit between a FreeSASA caller and the numeric hot path:
 *
 *   freesasa_calc_coord()      reference src/freesasa.h:475-479, src/freesasa.c:122-142
 *   freesasa_calc()            reference src/freesasa_internal.h:119-122, src/freesasa.c:76-120
 *   freesasa_lee_richards()    reference src/freesasa_internal.h:100-103, src/sasa_lr.c:156-216
 *   freesasa_shrake_rupley()   reference src/freesasa_internal.h:74-77,   src/sasa_sr.c:168-224
 *   freesasa_result_free()     reference src/freesasa.h:530, src/freesasa.c:68-74
 *   freesasa_default_parameters, freesasa_set_verbosity/get_verbosity, freesasa_set_err_out
 *
 * Same argument meaning, same return codes, same parameter validation and the same error/warning
 * text conventions ("freesasa: error: ...").  A program written against freesasa.h that only uses
 * these calls can link this library instead of libfreesasa; a full FreeSASA build swaps its
 * src/sasa_lr.c + src/sasa_sr.c + src/nb.c for the shim in INTEGRATION.md and keeps everything else.
 *
 * Around that path, the same library provides the host rows of SURVEY.md §8(f) under the reference's names, each
 * declaration below citing the lines it mirrors: classifiers (csrc/radii.c), structures and PDB reading (csrc/ingest.c),
 * the result tree, class sums and the PDB writer (csrc/areas.c), selections (csrc/select.c).  Results are bit-identical to
 * the reference's.
 *
 * Additive (not in the reference): freesasa_calc_coord_batch(), freesasa_calc_structure_batch(), freesasa_calc_tree_batch().
 */
#ifndef FREESASA_B200_HOST_H
#define FREESASA_B200_HOST_H

#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/freesasa.h:89-92 */
enum freesasa_algorithm { FREESASA_LEE_RICHARDS, FREESASA_SHRAKE_RUPLEY };
/* reference src/freesasa.h:104-109 */
enum freesasa_verbosity { FREESASA_V_NORMAL, FREESASA_V_NOWARNINGS, FREESASA_V_SILENT, FREESASA_V_DEBUG };
/* reference src/freesasa.h:151-155 */
enum freesasa_error_codes { FREESASA_SUCCESS = 0, FREESASA_FAIL = -1, FREESASA_WARN = -2 };

#ifndef __cplusplus
typedef enum freesasa_algorithm freesasa_algorithm;
typedef enum freesasa_verbosity freesasa_verbosity;
#endif

#define FREESASA_DEF_ALGORITHM FREESASA_LEE_RICHARDS /* src/freesasa.h:115-118 */
#define FREESASA_DEF_PROBE_RADIUS 1.4
#define FREESASA_DEF_SR_N 100
#define FREESASA_DEF_LR_N 20

/* reference src/freesasa.h:232-238 — layout must not change */
struct freesasa_parameters {
    freesasa_algorithm alg;
    double probe_radius;
    int shrake_rupley_n_points;
    int lee_richards_n_slices;
    int n_threads; /* validated as in the reference (<= 16), otherwise unused: the GPU grid replaces the pthread split */
};
/* reference src/freesasa.h:267-272 */
struct freesasa_result {
    double total;
    double *sasa; /* malloc'd; freed by freesasa_result_free() */
    int n_atoms;
    struct freesasa_parameters parameters;
};
/* reference src/coord.h:26-38 */
typedef struct coord_t {
    int n;
    int is_linked;
    double *xyz;
} coord_t;

#ifndef __cplusplus
typedef struct freesasa_parameters freesasa_parameters;
typedef struct freesasa_result freesasa_result;
#endif

/* ---- row f-1 of the scope table: classifiers and structure ingest ---------------------------------- */
/* reference src/freesasa.h:163-167 */
enum freesasa_atom_class { FREESASA_ATOM_APOLAR = 0, FREESASA_ATOM_POLAR = 1, FREESASA_ATOM_UNKNOWN = 2 };
/* reference src/freesasa.h:182-191 */
enum freesasa_structure_options {
    FREESASA_INCLUDE_HETATM = 1,
    FREESASA_INCLUDE_HYDROGEN = 1 << 2,
    FREESASA_SEPARATE_MODELS = 1 << 3,
    FREESASA_SEPARATE_CHAINS = 1 << 4,
    FREESASA_JOIN_MODELS = 1 << 5,
    FREESASA_HALT_AT_UNKNOWN = 1 << 6,
    FREESASA_SKIP_UNKNOWN = 1 << 7,
    FREESASA_RADIUS_FROM_OCCUPANCY = 1 << 8
};
/* reference src/freesasa.h:289-297 */
struct freesasa_nodearea {
    const char *name;
    double total, main_chain, side_chain, polar, apolar, unknown;
};
#define FREESASA_CONFLICTING_CLASSIFIERS "conflicting-classifiers" /* src/freesasa.h:133 */

/* reference src/freesasa.h:326-376: one _atom_site row of an mmCIF file as the reference's gemmi-based reader hands it over,
 * and a group of (up to three-character) chain labels */
struct freesasa_cif_atom {
    const char *group_PDB;
    const char auth_asym_id;
    const char *auth_seq_id, *pdbx_PDB_ins_code, *auth_comp_id, *auth_atom_id, *label_alt_id, *type_symbol;
    const double Cartn_x, Cartn_y, Cartn_z;
};
struct freesasa_cif_atom_lcl {
    const char *group_PDB, *auth_asym_id, *auth_seq_id, *pdbx_PDB_ins_code, *auth_comp_id, *auth_atom_id, *label_alt_id, *type_symbol;
    const double Cartn_x, Cartn_y, Cartn_z;
};
struct freesasa_chain_group {
    const char **chains;
    size_t n;
};
#ifndef __cplusplus
typedef struct freesasa_cif_atom freesasa_cif_atom;
typedef struct freesasa_cif_atom_lcl freesasa_cif_atom_lcl;
typedef struct freesasa_chain_group freesasa_chain_group;
#endif

typedef struct freesasa_classifier freesasa_classifier; /* opaque, src/freesasa.h:345 */
typedef struct freesasa_structure freesasa_structure;   /* opaque, src/freesasa.h:353 */
typedef struct freesasa_selection freesasa_selection;   /* opaque, src/freesasa.h:369 */
#ifndef __cplusplus
typedef enum freesasa_atom_class freesasa_atom_class;
typedef struct freesasa_nodearea freesasa_nodearea;
#endif

/* built-in classifiers, reference src/freesasa.h:418-436 (tables generated from the compiled reference) */
extern const freesasa_classifier freesasa_protor_classifier;
extern const freesasa_classifier freesasa_naccess_classifier;
extern const freesasa_classifier freesasa_oons_classifier;
#define freesasa_default_classifier freesasa_protor_classifier /* src/freesasa.h:124 */

/* reference src/freesasa.h:545-607, src/classifier.c:781-866 */
freesasa_classifier *freesasa_classifier_from_file(FILE *file);
void freesasa_classifier_free(freesasa_classifier *classifier);
double freesasa_classifier_radius(const freesasa_classifier *classifier, const char *res_name, const char *atom_name);
freesasa_atom_class freesasa_classifier_class(const freesasa_classifier *classifier, const char *res_name,
                                              const char *atom_name);
const char *freesasa_classifier_class2str(freesasa_atom_class atom_class);
const char *freesasa_classifier_name(const freesasa_classifier *classifier);
/* internal to the reference (src/classifier.h:72-78, src/freesasa_internal.h), exported for the parity tests */
double freesasa_guess_radius(const char *symbol);
const freesasa_nodearea *freesasa_classifier_residue_reference(const freesasa_classifier *classifier, const char *res_name);
int freesasa_atom_is_backbone(const char *atom_name);

/* reference src/freesasa.h:749-1446, src/structure.c */
freesasa_structure *freesasa_structure_new(void);
void freesasa_structure_free(freesasa_structure *structure);
freesasa_structure *freesasa_structure_from_pdb(FILE *pdb, const freesasa_classifier *classifier, int options);
freesasa_structure **freesasa_structure_array(FILE *pdb, int *n, const freesasa_classifier *classifier, int options);
int freesasa_structure_add_atom(freesasa_structure *structure, const char *atom_name, const char *residue_name,
                                const char *residue_number, char chain_label, double x, double y, double z);
int freesasa_structure_add_atom_wopt(freesasa_structure *structure, const char *atom_name, const char *residue_name,
                                     const char *residue_number, char chain_label, double x, double y, double z,
                                     const freesasa_classifier *classifier, int options);
freesasa_structure *freesasa_structure_get_chains(const freesasa_structure *structure, const char *chains,
                                                  const freesasa_classifier *classifier, int options);
freesasa_structure *freesasa_structure_get_chains_lcl(const freesasa_structure *structure, const freesasa_chain_group *chains,
                                                      const freesasa_classifier *classifier, int options);
int freesasa_structure_add_cif_atom(freesasa_structure *structure, freesasa_cif_atom *atom, const freesasa_classifier *classifier,
                                    int options);
int freesasa_structure_add_cif_atom_lcl(freesasa_structure *structure, freesasa_cif_atom_lcl *atom,
                                        const freesasa_classifier *classifier, int options);
int freesasa_structure_chain_atoms_lcl(const freesasa_structure *structure, const char *chain, int *first, int *last);
int freesasa_structure_chain_residues_lcl(const freesasa_structure *structure, const char *chain, int *first, int *last);
const char *freesasa_structure_residue_chain_lcl(const freesasa_structure *structure, int r_i);
const char *freesasa_structure_chain_labels(const freesasa_structure *structure);
int freesasa_structure_n(const freesasa_structure *structure);
int freesasa_structure_n_residues(const freesasa_structure *structure);
int freesasa_structure_n_chains(const freesasa_structure *structure);
const double *freesasa_structure_radius(const freesasa_structure *structure);
void freesasa_structure_set_radius(freesasa_structure *structure, const double *radii);
const char *freesasa_structure_atom_name(const freesasa_structure *structure, int i);
const char *freesasa_structure_atom_res_name(const freesasa_structure *structure, int i);
const char *freesasa_structure_atom_res_number(const freesasa_structure *structure, int i);
char freesasa_structure_atom_chain(const freesasa_structure *structure, int i);
const char *freesasa_structure_atom_chain_lcl(const freesasa_structure *structure, int i);
const char *freesasa_structure_atom_symbol(const freesasa_structure *structure, int i);
double freesasa_structure_atom_radius(const freesasa_structure *structure, int i);
void freesasa_structure_atom_set_radius(freesasa_structure *structure, int i, double radius);
freesasa_atom_class freesasa_structure_atom_class(const freesasa_structure *structure, int i);
const char *freesasa_structure_atom_pdb_line(const freesasa_structure *structure, int i);
const char *freesasa_structure_residue_name(const freesasa_structure *structure, int r_i);
const char *freesasa_structure_residue_number(const freesasa_structure *structure, int r_i);
char freesasa_structure_residue_chain(const freesasa_structure *structure, int r_i);
const freesasa_nodearea *freesasa_structure_residue_reference(const freesasa_structure *structure, int r_i);
int freesasa_structure_residue_atoms(const freesasa_structure *structure, int r_i, int *first, int *last);
int freesasa_structure_chain_atoms(const freesasa_structure *structure, char chain, int *first, int *last);
int freesasa_structure_chain_residues(const freesasa_structure *structure, char chain, int *first, int *last);
const char *freesasa_structure_chain_label(const freesasa_structure *structure, int index);
int freesasa_structure_model(const freesasa_structure *structure);
void freesasa_structure_set_model(freesasa_structure *structure, int model);
const char *freesasa_structure_classifier_name(const freesasa_structure *structure);
const double *freesasa_structure_coord_array(const freesasa_structure *structure);
const coord_t *freesasa_structure_xyz(const freesasa_structure *structure);
freesasa_result *freesasa_calc_structure(const freesasa_structure *structure, const freesasa_parameters *parameters);
/* Additive (row f-2): all structures of an array (NMR models, separated chains; the serial loop of
 * src/main.cc:334-362) in ONE device pass.  results[k] as in freesasa_calc_coord_batch(). */
int freesasa_calc_structure_batch(int n_struct, freesasa_structure *const *structures,
                                  const freesasa_parameters *parameters, freesasa_result **results);

/* ---- row f-3: areas of residues, chains and structures (the result tree) ------------------------------ */
/* reference src/freesasa.h:307-315 */
enum freesasa_nodetype {
    FREESASA_NODE_ATOM,
    FREESASA_NODE_RESIDUE,
    FREESASA_NODE_CHAIN,
    FREESASA_NODE_STRUCTURE,
    FREESASA_NODE_RESULT,
    FREESASA_NODE_ROOT,
    FREESASA_NODE_NONE
};
typedef struct freesasa_node freesasa_node; /* opaque, src/freesasa.h:361 */
#ifndef __cplusplus
typedef enum freesasa_nodetype freesasa_nodetype;
#endif
extern const freesasa_nodearea freesasa_nodearea_null; /* src/freesasa_internal.h, src/node.c:64 */

/* reference src/freesasa.h:497-519,1460-1549 */
freesasa_node *freesasa_calc_tree(const freesasa_structure *structure, const freesasa_parameters *parameters, const char *name);
freesasa_nodearea freesasa_result_classes(const freesasa_structure *structure, const freesasa_result *result);
/* Additive (row f-2): all structures in ONE device pass, then one tree each (built concurrently); trees[k] is what
 * freesasa_calc_tree(structures[k], parameters, names ? names[k] : NULL) returns.  FREESASA_SUCCESS or FREESASA_FAIL. */
int freesasa_calc_tree_batch(int n_struct, freesasa_structure *const *structures, const freesasa_parameters *parameters,
                             const char *const *names, freesasa_node **trees);
freesasa_node *freesasa_tree_new(void);
freesasa_node *freesasa_tree_init(const freesasa_result *result, const freesasa_structure *structure, const char *name);
int freesasa_tree_add_result(freesasa_node *tree, const freesasa_result *result, const freesasa_structure *structure,
                             const char *name);
int freesasa_tree_join(freesasa_node *tree1, freesasa_node **tree2);
int freesasa_node_free(freesasa_node *root);
/* reference src/freesasa.h:1559-1850 */
const freesasa_nodearea *freesasa_node_area(const freesasa_node *node);
freesasa_node *freesasa_node_children(freesasa_node *node);
freesasa_node *freesasa_node_next(freesasa_node *node);
freesasa_node *freesasa_node_parent(freesasa_node *node);
freesasa_nodetype freesasa_node_type(const freesasa_node *node);
const char *freesasa_node_name(const freesasa_node *node);
const char *freesasa_node_classified_by(const freesasa_node *node);
int freesasa_node_atom_is_polar(const freesasa_node *node);
int freesasa_node_atom_is_mainchain(const freesasa_node *node);
double freesasa_node_atom_radius(const freesasa_node *node);
const char *freesasa_node_atom_pdb_line(const freesasa_node *node);
const char *freesasa_node_atom_residue_number(const freesasa_node *node);
const char *freesasa_node_atom_residue_name(const freesasa_node *node);
const char *freesasa_node_atom_chain(const freesasa_node *node);
int freesasa_node_residue_n_atoms(const freesasa_node *node);
const char *freesasa_node_residue_number(const freesasa_node *node);
const freesasa_nodearea *freesasa_node_residue_reference(const freesasa_node *node);
int freesasa_node_chain_n_residues(const freesasa_node *node);
int freesasa_node_structure_n_chains(const freesasa_node *node);
int freesasa_node_structure_n_atoms(const freesasa_node *node);
int freesasa_node_structure_model(const freesasa_node *node);
const char *freesasa_node_structure_chain_labels(const freesasa_node *node);
const freesasa_result *freesasa_node_structure_result(const freesasa_node *node);
const freesasa_parameters *freesasa_node_result_parameters(const freesasa_node *node);
int freesasa_node_structure_add_selection(freesasa_node *node, const freesasa_selection *selection);
const freesasa_selection **freesasa_node_structure_selections(const freesasa_node *node);
/* row f-4, selections: reference src/freesasa.h:610-690,1855-1882, src/selection.c.  The command language is the
 * reference's ("name, resn ala+arg and not chain A", "s, resi 10-20+\-3", ...; doc/doxy-main.md "Selection syntax"). */
freesasa_selection *freesasa_selection_new(const char *command, const freesasa_structure *structure, const freesasa_result *result);
void freesasa_selection_free(freesasa_selection *selection);
freesasa_selection *freesasa_selection_clone(const freesasa_selection *selection);
const char *freesasa_selection_name(const freesasa_selection *selection);
const char *freesasa_selection_command(const freesasa_selection *selection);
double freesasa_selection_area(const freesasa_selection *selection);
int freesasa_selection_n_atoms(const freesasa_selection *selection);
int freesasa_select_area(const char *command, char *name, double *area, const freesasa_structure *structure,
                         const freesasa_result *result);
#define FREESASA_MAX_SELECTION_NAME 50
/* row f-4, per-atom writer: reference src/freesasa_internal.h:200, src/pdb.c:347-375 (what the CLI's --format=pdb and
 * freesasa_tree_export(..., FREESASA_PDB) emit) */
int freesasa_write_pdb(FILE *output, freesasa_node *root);
/* per-residue-type and per-residue listings, reference src/freesasa_internal.h:176-190, src/log.c:150-246 (the CLI's
 * --format=res / --format=seq, pinned by the reference's tests/data/restype.reference and seq.reference) */
int freesasa_write_res(FILE *log, freesasa_node *root);
int freesasa_write_seq(FILE *log, freesasa_node *root);
int freesasa_classify_n_residue_types(void);
int freesasa_classify_residue(const char *res_name);
const char *freesasa_classify_residue_name(int residue_type);
extern const char *freesasa_string;
/* reference src/freesasa_internal.h (used by the tree and the writers) */
int freesasa_atom_nodearea(freesasa_nodearea *area, const freesasa_structure *structure, const freesasa_result *result,
                           int atom_index);
void freesasa_add_nodearea(freesasa_nodearea *sum, const freesasa_nodearea *term);
void freesasa_range_nodearea(freesasa_nodearea *area, const freesasa_structure *structure, const freesasa_result *result,
                             int first_atom, int last_atom);

extern const freesasa_parameters freesasa_default_parameters;
extern const int FREESASA_DEF_NUMBER_THREADS;

freesasa_result *freesasa_calc_coord(const double *xyz, const double *radii, int n,
                                     const freesasa_parameters *parameters);
freesasa_result *freesasa_calc(const coord_t *c, const double *radii, const freesasa_parameters *parameters);
int freesasa_lee_richards(double *sasa, const coord_t *c, const double *radii, const freesasa_parameters *param);
int freesasa_shrake_rupley(double *sasa, const coord_t *c, const double *radii, const freesasa_parameters *param);
void freesasa_result_free(freesasa_result *result);

int freesasa_set_verbosity(freesasa_verbosity v);
freesasa_verbosity freesasa_get_verbosity(void);
void freesasa_set_err_out(FILE *err);
FILE *freesasa_get_err_out(void);

/* Additive: n_struct independent structures in one device pass.  results[k] is NULL-initialised by
 * the callee and receives a result per structure; returns FREESASA_SUCCESS or FREESASA_FAIL (in
 * which case no result is left allocated). */
int freesasa_calc_coord_batch(int n_struct, const double *const *xyz, const double *const *radii,
                              const int *n_atoms, const freesasa_parameters *parameters,
                              freesasa_result **results);

#ifdef __cplusplus
}
#endif
#endif
